"""Thin Python wrappers over the C ABI (include/vaegan_b200.h): tensors in, kernel launches out.

Everything here enqueues hand-written CUDA kernels on the current torch stream; torch is used only for
memory (allocation, views).  Every launching function is a ``torch.library`` custom op, registered under the name of the
C entry point it wraps (``torch.ops.vaegan.vg_norm_apply`` ...; see dispatch.py): calling ``ops.norm_apply(...)`` goes
through the dispatcher.  NHWC bf16 activations are plain torch tensors of shape [N, H, W, C] whose last
stride is 1 and whose pixel stride ``ld = stride(2)`` may exceed C (a channel slice of a wider buffer).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from .dispatch import launch_op

BF16 = torch.bfloat16
F32 = torch.float32
_LL5 = C.c_longlong * 5

# Activation storage of the package: "bf16" (default; bf16 storage + bf16 tensor-core inputs, fp32 accumulation) or
# "fp32" (high-accuracy mode: fp32 storage, every tensor-core operand split into three bf16 planes so that products
# carry ~24 mantissa bits; 6x the tensor-core work, used for the fp32-tolerance parity runs).
_ACT_DTYPE = BF16


def set_precision(mode: str) -> None:
    global _ACT_DTYPE
    assert mode in ("bf16", "fp32"), mode
    _ACT_DTYPE = BF16 if mode == "bf16" else F32
    # the stock recurrent text encoder runs through cuDNN, which defaults to TF32 tensor cores for fp32 RNNs/convs:
    # in the high-accuracy mode that 1e-3 error would dominate everything downstream
    torch.backends.cudnn.allow_tf32 = mode == "bf16"


def act_dtype() -> torch.dtype:
    return _ACT_DTYPE


def dcode(t: torch.Tensor) -> int:
    """dtype code of the C ABI for an activation tensor."""
    return 0 if t.dtype == BF16 else 1


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def nhwc_ok(t: torch.Tensor) -> bool:
    n, h, w, c = t.shape
    ld = t.stride(2)
    return (t.dim() == 4 and t.stride(3) == 1 and ld >= c and (h == 1 or t.stride(1) == w * ld) and
            (n == 1 or t.stride(0) == h * w * ld) and t.data_ptr() % 16 == 0 and ld % 8 == 0)


def dense_nhwc(t: torch.Tensor, dtype=None) -> torch.Tensor:
    out = torch.empty(t.shape, dtype=dtype or t.dtype, device=t.device)
    strided_copy(t, out)
    return out


@launch_op("vg_split3(Tensor t) -> Tensor")
def split3(t: torch.Tensor) -> torch.Tensor:
    """fp32 NHWC view [N,H,W,C] -> bf16 [N,H,W,3*Cp] holding the hi | mid | lo planes (Cp = C rounded up to 64)."""
    assert t.dtype == F32 and t.stride(3) == 1
    n, h, w, c = t.shape
    cp = round_up(c, 64)
    assert (h == 1 or t.stride(1) == w * t.stride(2)) and (n == 1 or t.stride(0) == h * w * t.stride(2))
    out = torch.empty((n, h, w, 3 * cp), dtype=BF16, device=t.device)
    _lib.call("vg_split3", _p(t), t.stride(2), C.c_longlong(n * h * w), c, cp, _p(out), stream())
    return out


def ld_of(t: torch.Tensor) -> int:
    return t.stride(2)


# ----------------------------------------------------------------------------------------------
# data movement
# ----------------------------------------------------------------------------------------------
@launch_op("vg_strided_copy(Tensor src, Tensor(a!) dst, Tensor? scale=None, bool scale_inverse=False, bool accumulate=False) -> ()")
def strided_copy(src: torch.Tensor, dst: torch.Tensor, scale: Optional[torch.Tensor] = None,
                 scale_inverse: bool = False, accumulate: bool = False) -> None:
    """dst[...] (=|+=) scale * src[...] for two tensors of equal shape (<= 5 dims), any strides, fp32/bf16."""
    assert src.shape == dst.shape and src.dim() <= 5, (src.shape, dst.shape)
    if src.numel() == 0:
        return
    pad = 5 - src.dim()
    dims = [1] * pad + list(src.shape)
    iss = [0] * pad + list(src.stride())
    oss = [0] * pad + list(dst.stride())
    code = {F32: 0, BF16: 1}
    _lib.call("vg_strided_copy", _p(src), code[src.dtype], _p(dst), code[dst.dtype], _LL5(*dims), _LL5(*iss), _LL5(*oss),
              _p(scale), int(scale_inverse), int(accumulate), stream())


# ----------------------------------------------------------------------------------------------
# normalisation + activation (+ pool)
# ----------------------------------------------------------------------------------------------
@launch_op("vg_norm_stats(Tensor x, bool per_sample) -> Tensor")
def norm_stats(x: torch.Tensor, per_sample: bool) -> torch.Tensor:
    n, h, w, c = x.shape
    groups = n if per_sample else 1
    rows = h * w if per_sample else n * h * w
    sums = torch.empty(groups, 2, c, dtype=F32, device=x.device)
    _lib.call("vg_norm_stats", _p(x), ld_of(x), 0, groups, C.c_longlong(rows), c, _p(sums), dcode(x), stream())
    return sums


@launch_op("vg_norm_stats_rows(Tensor x, int virt_h) -> Tensor")
def norm_stats_rows(x: torch.Tensor, virt_h: int) -> torch.Tensor:
    """Batch statistics of a tensor whose h rows stand for ``virt_h`` rows (interior rows all equal)."""
    n, h, w, c = x.shape
    sums = torch.empty(1, 2, c, dtype=F32, device=x.device)
    assert x.is_contiguous()
    _lib.call("vg_norm_stats_rows", _p(x), ld_of(x), 0, n, h, w, c, virt_h, _p(sums), dcode(x), stream())
    return sums


@launch_op("vg_norm_finalize(Tensor sums, int rows, float eps, float momentum=0.1, Tensor(a!)? running_mean=None, Tensor(b!)? running_var=None, Tensor(c!)? num_batches_tracked=None) -> Tensor")
def norm_finalize(sums: torch.Tensor, rows: int, eps: float, momentum: float = 0.1,
                  running_mean: Optional[torch.Tensor] = None, running_var: Optional[torch.Tensor] = None,
                  num_batches_tracked: Optional[torch.Tensor] = None) -> torch.Tensor:
    groups, _, c = sums.shape
    mr = torch.empty_like(sums)
    _lib.call("vg_norm_finalize", _p(sums), groups, C.c_longlong(rows), c, C.c_float(eps), _p(mr), C.c_float(momentum),
              _p(running_mean), _p(running_var), _p(num_batches_tracked), stream())
    return mr


@launch_op("vg_norm_apply(Tensor x, Tensor mean_rstd, Tensor? gamma, Tensor? beta, int act, Tensor(a!) y, Tensor(b!)? pool=None) -> ()")
def norm_apply(x: torch.Tensor, mean_rstd: torch.Tensor, gamma, beta, act: int, y: torch.Tensor,
               pool: Optional[torch.Tensor] = None) -> None:
    n, h, w, c = x.shape
    d = _lib.VgNormApply()
    d.x, d.x_ld, d.x_coff = x.data_ptr(), ld_of(x), 0
    d.n, d.h, d.w, d.c = n, h, w, c
    d.mean_rstd, d.per_sample = mean_rstd.data_ptr(), int(mean_rstd.shape[0] > 1 or False)
    d.gamma = gamma.data_ptr() if gamma is not None else None
    d.beta = beta.data_ptr() if beta is not None else None
    d.act = act
    d.y, d.y_ld, d.y_coff = y.data_ptr(), ld_of(y), 0
    if pool is not None:
        d.pool, d.p_ld, d.p_coff = pool.data_ptr(), ld_of(pool), 0
    d.dtype = dcode(x)
    assert y.dtype == x.dtype
    _lib.call("vg_norm_apply", C.byref(d), stream())


@launch_op("vg_norm_backward(Tensor x, Tensor? dy, Tensor? dpool, Tensor mean_rstd, bool per_sample, Tensor? gamma, Tensor? beta, int act, Tensor(a!) dx, Tensor(b!)? dgamma, Tensor(c!)? dbeta, bool accumulate=False, int virt_h=0) -> ()")
def norm_backward(x, dy, dpool, mean_rstd, per_sample, gamma, beta, act, dx, dgamma, dbeta, accumulate=False,
                  virt_h: int = 0):
    n, h, w, c = x.shape
    groups = n if per_sample else 1
    sums = torch.empty(groups, 2, c, dtype=F32, device=x.device)
    d = _lib.VgNormBackward()
    d.x, d.x_ld, d.x_coff = x.data_ptr(), ld_of(x), 0
    if dy is not None:
        d.dy, d.dy_ld, d.dy_coff = dy.data_ptr(), ld_of(dy), 0
    if dpool is not None:
        d.dpool, d.dp_ld, d.dp_coff = dpool.data_ptr(), ld_of(dpool), 0
    d.n, d.h, d.w, d.c = n, h, w, c
    d.mean_rstd, d.per_sample = mean_rstd.data_ptr(), int(per_sample)
    d.gamma = gamma.data_ptr() if gamma is not None else None
    d.beta = beta.data_ptr() if beta is not None else None
    d.act = act
    d.sums = sums.data_ptr()
    d.dx, d.dx_ld, d.dx_coff = dx.data_ptr(), ld_of(dx), 0
    d.dgamma = dgamma.data_ptr() if dgamma is not None else None
    d.dbeta = dbeta.data_ptr() if dbeta is not None else None
    d.accumulate = int(accumulate)
    d.dtype = dcode(x)
    d.virt_h = virt_h
    assert virt_h == 0 or (x.is_contiguous() and dy.is_contiguous() and dx.is_contiguous())
    assert dx.dtype == x.dtype and (dy is None or dy.dtype == x.dtype) and (dpool is None or dpool.dtype == x.dtype)
    _lib.call("vg_norm_backward", C.byref(d), stream())


@launch_op("vg_maxpool2x2_fwd(Tensor x, Tensor(a!) y) -> ()")
def maxpool_fwd(x, y):
    n, h, w, c = x.shape
    assert x.dtype == y.dtype and y.is_contiguous() and nhwc_ok(x)
    _lib.call("vg_maxpool2x2_fwd", _p(x), ld_of(x), _p(y), n, h, w, c, dcode(x), stream())


@launch_op("vg_maxpool2x2_bwd(Tensor x, Tensor dy, Tensor(a!) dx) -> ()")
def maxpool_bwd(x, dy, dx):
    n, h, w, c = x.shape
    assert x.dtype == dy.dtype == dx.dtype and dy.is_contiguous() and dx.is_contiguous() and nhwc_ok(x)
    _lib.call("vg_maxpool2x2_bwd", _p(x), ld_of(x), _p(dy), _p(dx), n, h, w, c, dcode(x), stream())


@launch_op("vg_act_bwd(Tensor y, Tensor dy, Tensor(a!) dx, int act) -> ()")
def act_bwd(y, dy, dx, act):
    n, h, w, c = y.shape
    assert y.dtype == dy.dtype == dx.dtype
    _lib.call("vg_act_bwd", _p(y), ld_of(y), _p(dy), ld_of(dy), _p(dx), ld_of(dx), C.c_longlong(n * h * w), c, act,
              dcode(y), stream())


@launch_op("vg_act_fwd(Tensor(a!) y, int act) -> ()")
def act_fwd_(y, act):
    n, h, w, c = y.shape
    _lib.call("vg_act_fwd", _p(y), ld_of(y), C.c_longlong(n * h * w), c, act, dcode(y), stream())


@launch_op("vg_colsum_f32(Tensor t2d, Tensor(a!) out, bool accumulate=False) -> ()")
def colsum_f32(t2d, out, accumulate=False):
    rows, cols = t2d.shape
    _lib.call("vg_colsum_f32", _p(t2d), C.c_longlong(rows), cols, t2d.stride(0), _p(out), int(accumulate), stream())


# ----------------------------------------------------------------------------------------------
# FiLM / upsample / im2col / small-N convs
# ----------------------------------------------------------------------------------------------
@launch_op("vg_film_fwd(Tensor gb, Tensor x, Tensor(a!) y) -> ()")
def film_fwd(gb, x, y):
    n, h, w, c = x.shape
    assert gb.dtype == x.dtype == y.dtype
    _lib.call("vg_film_fwd", _p(gb), _p(x), ld_of(x), 0, _p(y), C.c_longlong(n * h * w), c, dcode(x), stream())


@launch_op("vg_film_bwd(Tensor gb, Tensor x, Tensor dy, Tensor(a!) dgb, Tensor(b!) dx) -> ()")
def film_bwd(gb, x, dy, dgb, dx):
    n, h, w, c = x.shape
    assert gb.dtype == x.dtype == dy.dtype == dgb.dtype == dx.dtype
    _lib.call("vg_film_bwd", _p(gb), _p(x), ld_of(x), 0, _p(dy), _p(dgb), _p(dx), ld_of(dx), 0,
              C.c_longlong(n * h * w), c, dcode(x), stream())


@launch_op("vg_film_rows_fwd(Tensor gb3, Tensor x, Tensor(a!) y) -> ()")
def film_rows_fwd(gb3, x, y):
    n, h, w, c = x.shape
    assert gb3.dtype == x.dtype == y.dtype and gb3.is_contiguous() and y.is_contiguous()
    assert tuple(gb3.shape) == (n, 3, w, 2 * c), (gb3.shape, x.shape)
    _lib.call("vg_film_rows_fwd", _p(gb3), 3, _p(x), ld_of(x), 0, _p(y), n, h, w, c, dcode(x), stream())


@launch_op("vg_film_rows_bwd(Tensor gb3, Tensor x, Tensor dy, Tensor(a!) dgb3, Tensor(b!) dx) -> ()")
def film_rows_bwd(gb3, x, dy, dgb3, dx):
    n, h, w, c = x.shape
    assert gb3.dtype == x.dtype == dy.dtype == dgb3.dtype == dx.dtype
    assert gb3.is_contiguous() and dy.is_contiguous() and dgb3.is_contiguous()
    _lib.call("vg_film_rows_bwd", _p(gb3), 3, _p(x), ld_of(x), 0, _p(dy), _p(dgb3), _p(dx), ld_of(dx), 0, n, h, w, c,
              dcode(x), stream())


@launch_op("vg_upsample_w_fwd(Tensor t, Tensor(a!) y) -> ()")
def upsample_w_fwd(t, y):
    n, _, w0, c = t.shape
    _, h, w, _ = y.shape
    assert t.dtype == y.dtype
    _lib.call("vg_upsample_w_fwd", _p(t), ld_of(t), 0, n, w0, c, _p(y), h, w, dcode(t), stream())


@launch_op("vg_upsample_w_bwd(Tensor dy, Tensor(a!) dt) -> ()")
def upsample_w_bwd(dy, dt):
    n, h, w, c = dy.shape
    w0 = dt.shape[2]
    _lib.call("vg_upsample_w_bwd", _p(dy), n, h, w, c, w0, _p(dt), dcode(dy), stream())


@launch_op("vg_upsample_h_fwd(Tensor t, Tensor(a!) y) -> ()")
def upsample_h_fwd(t, y):
    """Bilinear resize along H only: t [n,h0,w,c] -> y [n,h,w,c], both dense."""
    n, h0, w, c = t.shape
    assert t.dtype == y.dtype and t.is_contiguous() and y.is_contiguous() and tuple(y.shape) == (n, y.shape[1], w, c)
    _lib.call("vg_upsample_h_fwd", _p(t), n, h0, w, c, _p(y), y.shape[1], dcode(t), stream())


@launch_op("vg_upsample_h_bwd(Tensor dy, Tensor(a!) dt) -> ()")
def upsample_h_bwd(dy, dt):
    """dy [n,h,w,c] dense (bf16 or fp32) -> dt fp32 [n,h0,w,c] (fully written)."""
    n, h, w, c = dy.shape
    assert dt.dtype == F32 and dy.is_contiguous() and dt.is_contiguous() and tuple(dt.shape) == (n, dt.shape[1], w, c)
    _lib.call("vg_upsample_h_bwd", _p(dy), n, h, w, c, dt.shape[1], _p(dt), dcode(dy), stream())


@launch_op("vg_channel_scale_fwd(Tensor x, Tensor scale, Tensor(a!) y) -> ()")
def channel_scale_fwd(x, scale, y):
    """y[..., ch] = x[..., ch] * scale[ch]; x, y NHWC views (y may be a channel slice), scale fp32 [c]."""
    n, h, w, c = x.shape
    assert x.dtype == y.dtype and scale.dtype == F32 and scale.is_contiguous() and scale.numel() == c
    assert nhwc_ok(x) and nhwc_ok(y) and tuple(y.shape) == tuple(x.shape)
    _lib.call("vg_channel_scale_fwd", _p(x), ld_of(x), _p(scale), _p(y), ld_of(y), 0, C.c_longlong(n * h * w), c,
              dcode(x), stream())


@launch_op("vg_channel_scale_bwd(Tensor x, Tensor dy, Tensor scale, Tensor(a!) dx, Tensor(b!) dscale) -> ()")
def channel_scale_bwd(x, dy, scale, dx, dscale):
    """dx = dy * scale, dscale[:c] = sum over pixels of dy * x (dscale: fp32 [2c], second half scratch)."""
    n, h, w, c = x.shape
    assert x.dtype == dy.dtype == dx.dtype and dscale.dtype == F32 and dscale.numel() == 2 * c
    assert nhwc_ok(x) and nhwc_ok(dy) and nhwc_ok(dx)
    _lib.call("vg_channel_scale_bwd", _p(x), ld_of(x), _p(dy), ld_of(dy), 0, _p(scale), _p(dx), ld_of(dx),
              C.c_longlong(n * h * w), c, _p(dscale), dcode(x), stream())


@launch_op("vg_im2col(Tensor src, int c, int kh, int kw, int stride, int pad, Tensor(a!) col) -> ()")
def im2col(src, c, kh, kw, stride, pad, col):
    n, h, w, _ = src.shape
    assert src.dtype == col.dtype
    _lib.call("vg_im2col", _p(src), n, h, w, ld_of(src), c, kh, kw, stride, pad, _p(col), col.shape[-1], dcode(src),
              stream())


@launch_op("vg_col2im(Tensor dcol, int n, int h, int w, int c, int kh, int kw, int stride, int pad, Tensor(a!) dsrc_nchw) -> ()")
def col2im(dcol, n, h, w, c, kh, kw, stride, pad, dsrc_nchw):
    _lib.call("vg_col2im", _p(dcol), dcol.shape[-1], n, h, w, c, kh, kw, stride, pad, _p(dsrc_nchw), dcode(dcol),
              stream())


@launch_op("vg_conv_smalln_fwd(Tensor x, Tensor wt, Tensor? bias, int kh, int kw, int pad, Tensor(a!) out) -> ()")
def smalln_fwd(x, wt, bias, kh, kw, pad, out):
    n, h, w, cin = x.shape
    _lib.call("vg_conv_smalln_fwd", _p(x), ld_of(x), 0, n, h, w, cin, _p(wt), _p(bias), wt.shape[0], kh, kw, pad,
              _p(out), dcode(x), stream())


@launch_op("vg_conv_smalln_dgrad(Tensor dy, Tensor wt, int kh, int kw, int pad, Tensor(a!) dx) -> ()")
def smalln_dgrad(dy, wt, kh, kw, pad, dx):
    n, h, w, cin = dx.shape
    _lib.call("vg_conv_smalln_dgrad", _p(dy), n, h, w, cin, _p(wt), wt.shape[0], kh, kw, pad, _p(dx), ld_of(dx), 0,
              dcode(dx), stream())


@launch_op("vg_conv_smalln_wgrad(Tensor dy, Tensor x, int kh, int kw, int pad, Tensor(a!) dw, Tensor(b!)? dbias) -> ()")
def smalln_wgrad(dy, x, kh, kw, pad, dw, dbias):
    n, h, w, cin = x.shape
    _lib.call("vg_conv_smalln_wgrad", _p(dy), _p(x), ld_of(x), 0, n, h, w, cin, dw.shape[0], kh, kw, pad, _p(dw),
              _p(dbias), dcode(x), stream())


# ----------------------------------------------------------------------------------------------
# losses / reparam / spectral norm / optimiser
# ----------------------------------------------------------------------------------------------
@launch_op("vg_reparam_kl_fwd(Tensor heads, Tensor bias_mu, Tensor bias_lv, Tensor eps) -> (Tensor, Tensor, Tensor, Tensor)")
def reparam_kl_fwd(heads, bias_mu, bias_lv, eps):
    b, z2 = heads.shape
    z = z2 // 2
    mu, lv, zo = (torch.empty(b, z, dtype=F32, device=heads.device) for _ in range(3))
    kl = torch.empty((), dtype=F32, device=heads.device)
    _lib.call("vg_reparam_kl_fwd", _p(heads), _p(bias_mu), _p(bias_lv), _p(eps), b, z, _p(mu), _p(lv), _p(zo), _p(kl),
              stream())
    return mu, lv, zo, kl


@launch_op("vg_reparam_kl_bwd(Tensor mu, Tensor lv, Tensor eps, Tensor? dz, Tensor? dmu, Tensor? dlv, Tensor? dkl, Tensor(a!)? dheads_bf16=None) -> Tensor")
def reparam_kl_bwd(mu, lv, eps, dz, dmu, dlv, dkl, dheads_bf16=None):
    b, z = mu.shape
    dheads = torch.empty(b, 2 * z, dtype=F32, device=mu.device)
    _lib.call("vg_reparam_kl_bwd", _p(mu), _p(lv), _p(eps), _p(dz), _p(dmu), _p(dlv), _p(dkl), b, z, _p(dheads),
              _p(dheads_bf16), dheads_bf16.stride(0) if dheads_bf16 is not None else 0, stream())
    return dheads


@launch_op("vg_sigmoid_fwd(Tensor pre_nhwc, Tensor(a!) y_nchw) -> ()")
def sigmoid_fwd(pre_nhwc, y_nchw):
    n, c, h, w = y_nchw.shape
    _lib.call("vg_sigmoid_fwd", _p(pre_nhwc), n, c, h * w, _p(y_nchw), stream())


@launch_op("vg_sigmoid_bwd(Tensor y_nchw, Tensor dy_nchw, Tensor(a!) dpre_nhwc) -> ()")
def sigmoid_bwd(y_nchw, dy_nchw, dpre_nhwc):
    n, c, h, w = y_nchw.shape
    _lib.call("vg_sigmoid_bwd", _p(y_nchw), _p(dy_nchw), n, c, h * w, _p(dpre_nhwc), stream())


@launch_op("vg_l1_fwd(Tensor a, Tensor b) -> Tensor")
def l1_fwd(a, b):
    out = torch.empty((), dtype=F32, device=a.device)
    _lib.call("vg_l1_fwd", _p(a), _p(b), C.c_longlong(a.numel()), _p(out), stream())
    return out


@launch_op("vg_l1_bwd(Tensor a, Tensor b, Tensor gout, Tensor(a!) da, bool accumulate=False) -> ()")
def l1_bwd(a, b, gout, da, accumulate=False):
    _lib.call("vg_l1_bwd", _p(a), _p(b), C.c_longlong(a.numel()), _p(gout), _p(da), int(accumulate), stream())


@launch_op("vg_hinge_fwd(Tensor p, int mode) -> Tensor")
def hinge_fwd(p, mode):
    out = torch.empty((), dtype=F32, device=p.device)
    _lib.call("vg_hinge_fwd", _p(p), C.c_longlong(p.numel()), mode, _p(out), stream())
    return out


@launch_op("vg_hinge_bwd(Tensor p, int mode, Tensor gout, Tensor(a!) dp) -> ()")
def hinge_bwd(p, mode, gout, dp):
    _lib.call("vg_hinge_bwd", _p(p), C.c_longlong(p.numel()), mode, _p(gout), _p(dp), stream())


@launch_op("vg_spectral_sigma(Tensor w2d, Tensor(a!) u, Tensor(b!) v, bool training, float eps=1e-12) -> Tensor")
def spectral_sigma(w2d, u, v, training, eps=1e-12):
    rows, cols = w2d.shape
    sigma = torch.empty((), dtype=F32, device=w2d.device)
    scratch = torch.empty(rows + cols + 2, dtype=F32, device=w2d.device)
    _lib.call("vg_spectral_sigma", _p(w2d), rows, cols, _p(u), _p(v), int(training), C.c_float(eps), _p(sigma),
              _p(scratch), stream())
    return sigma


@launch_op("vg_spectral_bwd(Tensor g2d, Tensor w2d, Tensor u, Tensor v, Tensor sigma, Tensor(a!) dw, bool accumulate=False) -> ()")
def spectral_bwd(g2d, w2d, u, v, sigma, dw, accumulate=False):
    rows, cols = w2d.shape
    scratch = torch.empty(1, dtype=F32, device=w2d.device)
    _lib.call("vg_spectral_bwd", _p(g2d), _p(w2d), _p(u), _p(v), _p(sigma), rows, cols, _p(dw), int(accumulate),
              _p(scratch), stream())


@launch_op("vg_gru_seq_fwd(Tensor xproj, Tensor w_hh, Tensor b_hh, Tensor(a!) out, Tensor(b!) gates) -> ()")
def gru_seq_fwd(xproj, w_hh, b_hh, out, gates):
    """Time recurrence of one bidirectional GRU layer; xproj [B,T,2,3H], out [B,T,2H], gates [2,B,T,4,H] (fp32)."""
    b, t = out.shape[0], out.shape[1]
    h = w_hh.shape[2]
    for x in (xproj, w_hh, b_hh, out, gates):
        assert x.dtype == F32 and x.is_contiguous()
    _lib.call("vg_gru_seq_fwd", _p(xproj), _p(w_hh), _p(b_hh), _p(out), _p(gates), b, t, h, stream())


@launch_op("vg_gru_seq_bwd(Tensor dout, Tensor out, Tensor gates, Tensor w_hh, Tensor(a!) dgx, Tensor(b!) dgh) -> ()")
def gru_seq_bwd(dout, out, gates, w_hh, dgx, dgh):
    b, t = out.shape[0], out.shape[1]
    h = w_hh.shape[2]
    for x in (dout, out, gates, w_hh, dgx, dgh):
        assert x.dtype == F32 and x.is_contiguous()
    _lib.call("vg_gru_seq_bwd", _p(dout), _p(out), _p(gates), _p(w_hh), _p(dgx), _p(dgh), b, t, h, stream())


# ----------------------------------------------------------------------------------------------
# text front end: tokenisation, embedding, sequence pooling
# ----------------------------------------------------------------------------------------------
@launch_op("vg_tokenize(Tensor codepoints, Tensor lut) -> Tensor")
def tokenize(codepoints: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    """codepoints: int32 device tensor holding uint32 UTF-32 code units (any shape); lut: int32 [lut_size].  -> int64 indices."""
    assert codepoints.dtype == torch.int32 and lut.dtype == torch.int32 and codepoints.is_contiguous() and lut.is_contiguous()
    idx = torch.empty(codepoints.shape, dtype=torch.long, device=codepoints.device)
    _lib.call("vg_tokenize", _p(codepoints), C.c_longlong(codepoints.numel()), _p(lut), lut.numel(), _p(idx), stream())
    return idx


@launch_op("vg_embedding_fwd(Tensor idx, Tensor weight) -> Tensor")
def embedding_fwd(idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    assert idx.dtype == torch.long and idx.is_contiguous() and weight.dtype == F32 and weight.is_contiguous()
    out = torch.empty(tuple(idx.shape) + (weight.shape[1],), dtype=F32, device=weight.device)
    _lib.call("vg_embedding_fwd", _p(idx), C.c_longlong(idx.numel()), _p(weight), weight.shape[0], weight.shape[1], _p(out),
              stream())
    return out


@launch_op("vg_embedding_bwd(Tensor idx, Tensor g, int vocab, int padding_idx) -> Tensor")
def embedding_bwd(idx: torch.Tensor, g: torch.Tensor, vocab: int, padding_idx: int) -> torch.Tensor:
    g = g.contiguous()
    assert g.dtype == F32
    dw = torch.empty((vocab, g.shape[-1]), dtype=F32, device=g.device)
    _lib.call("vg_embedding_bwd", _p(idx), C.c_longlong(idx.numel()), _p(g), vocab, g.shape[-1], padding_idx, _p(dw), stream())
    return dw


@launch_op("vg_seqpool_fwd(Tensor seq, Tensor(a!) out) -> ()")
def seqpool_fwd(seq: torch.Tensor, out: torch.Tensor) -> None:
    """seq: [b,l,c] (or NHWC [b,1,l,c]) rows of stride seq.stride(-2); out: NHWC [b,1,w,c]."""
    b, l, c = seq.shape[0], seq.shape[-2], seq.shape[-1]
    w = out.shape[2]
    assert seq.stride(-1) == 1 and out.stride(3) == 1 and seq.stride(0) == l * seq.stride(-2) and out.stride(0) == w * out.stride(2)
    _lib.call("vg_seqpool_fwd", _p(seq), dcode(seq), seq.stride(-2), b, l, c, w, _p(out), dcode(out), out.stride(2), stream())


@launch_op("vg_seqpool_bwd(Tensor dy, Tensor(a!) dseq) -> ()")
def seqpool_bwd(dy: torch.Tensor, dseq: torch.Tensor) -> None:
    """dy: NHWC [b,1,w,c]; dseq: [b,l,c] (or NHWC [b,1,l,c]), fully written."""
    b, l, c = dseq.shape[0], dseq.shape[-2], dseq.shape[-1]
    w = dy.shape[2]
    assert dy.stride(3) == 1 and dseq.stride(-1) == 1 and dy.stride(0) == w * dy.stride(2) and dseq.stride(0) == l * dseq.stride(-2)
    _lib.call("vg_seqpool_bwd", _p(dy), dcode(dy), dy.stride(2), b, l, c, w, _p(dseq), dcode(dseq), dseq.stride(-2), stream())


@launch_op("vg_sumsq(Tensor g, Tensor(a!) out, bool zero_first=True) -> ()")
def sumsq(g, out, zero_first=True):
    _lib.call("vg_sumsq", _p(g), C.c_longlong(g.numel()), _p(out), int(zero_first), stream())


@launch_op("vg_adam_step(Tensor(a!) p, Tensor(b!) g, Tensor(c!) m, Tensor(d!) v, float lr, float beta1, float beta2, float eps, int step, Tensor? gnorm_sq=None, float max_norm=0.0, bool write_back_grad=False) -> ()")
def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, gnorm_sq=None, max_norm=0.0, write_back_grad=False):
    _lib.call("vg_adam_step", _p(p), _p(g), _p(m), _p(v), C.c_longlong(p.numel()), C.c_float(lr), C.c_float(beta1),
              C.c_float(beta2), C.c_float(eps), int(step), _p(gnorm_sq), C.c_float(max_norm), int(write_back_grad),
              stream())
