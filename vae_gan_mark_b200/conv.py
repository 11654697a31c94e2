"""Tensor-core convolutions: geometry -> tap tables -> ``vg_conv_fprop`` / ``vg_conv_wgrad`` launches.

Activations are NHWC bf16 torch views ([N, H, W, C], last stride 1, pixel stride ``ld = stride(2)`` >= C so a
tensor may be a channel slice of a wider concat buffer).  ``ConvLinear`` describes one Conv2d-shaped linear
map and provides its three primitives (forward, data gradient, weight gradient); a ConvTranspose2d is the
adjoint of the Conv2d with the same weight tensor, so its forward is that conv's data gradient, its data
gradient is that conv's forward and its weight gradient is that conv's weight gradient with the operands
swapped (IOHW of the transpose == OIHW of its adjoint).  Reference layers: vae-gan.py:52-60,76-81,153-157;
vae-gan-v2.py:123-127,168-176,199-241; vae-gan-unet.py:148-154,194-221.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from ._lib import VgConvFprop, VgConvWgrad
from .ops import BF16, F32, round_up

Tap = Tuple[int, int, int, int]   # (c_base, dw, sh, dh)

# When set to a list, every tensor-core launch appends (kind, shape-key, flops, start_event, end_event): bench.py
# uses it to time the dominant kernel live on the launching stream.
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(e0, kind, key, flops):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILE.append((kind, key, flops, e0, e1))


def conv_taps(kh: int, kw: int, stride: int, ph: int, pw: int, ld: int) -> List[Tap]:
    """Taps of a (kh x kw, stride, pad) convolution reading its input through the stride view."""
    taps = []
    for r in range(kh):
        for q in range(kw):
            rr, qq = r - ph, q - pw
            taps.append((0, qq, 0, rr) if stride == 1 else ((qq % 2) * ld, qq // 2, rr % 2, rr // 2))
    return taps


def _fill_taps(dst, taps: Sequence[Tap]):
    assert len(taps) <= _lib.VG_MAX_TAPS, f"{len(taps)} taps > {_lib.VG_MAX_TAPS}"
    for i, t in enumerate(taps):
        for j in range(4):
            dst[i][j] = int(t[j])


def _chk(t: torch.Tensor, what: str):
    assert t.dtype == BF16 and ops.nhwc_ok(t), f"{what}: not an NHWC bf16 view {tuple(t.shape)} {t.stride()}"


def fprop(x: torch.Tensor, taps: Sequence[Tap], x_stride: int, cin: int, w: torch.Tensor, n_gemm: int,
          m: Tuple[int, int, int], out: torch.Tensor, out_kind: int = 0, su: Tuple[int, int] = (1, 1),
          sub0: Tuple[int, int] = (0, 0), cout_per_sub: Optional[int] = None, bias: Optional[torch.Tensor] = None,
          act: int = 0, ksplit: int = 0, force_bn: int = 0) -> None:
    """out[pixel, n] = sum_{tap, c<cin} x[pixel@tap, c] * w[n, tap*cin + c].  ``out`` is an NHWC view
    ([N, OH, OW, C']); ``w`` is bf16 [n_gemm, len(taps)*cin]."""
    _chk(x, "fprop x")
    assert w.dtype == BF16 and w.stride(1) == 1 and w.shape[1] == len(taps) * cin, (w.shape, len(taps), cin)
    assert out.stride(3) == 1
    d = VgConvFprop()
    d.x, d.x_n, d.x_h, d.x_w, d.x_ld, d.x_stride = x.data_ptr(), x.shape[0], x.shape[1], x.shape[2], x.stride(2), x_stride
    d.m_n, d.m_h, d.m_w = m
    d.cin, d.num_taps = cin, len(taps)
    _fill_taps(d.taps, taps)
    d.w, d.w_ld, d.n_gemm = w.data_ptr(), w.stride(0), n_gemm
    d.out, d.out_kind = out.data_ptr(), out_kind
    d.out_h, d.out_w, d.out_ld, d.out_coff = out.shape[1], out.shape[2], out.stride(2), 0
    d.su_h, d.su_w = su
    d.sub_h0, d.sub_w0 = sub0
    d.cout_per_sub = cout_per_sub or n_gemm
    d.bias = bias.data_ptr() if bias is not None else None
    d.act, d.ksplit, d.force_bn = act, ksplit, force_bn
    e0 = _prof_begin()
    _lib.call("vg_conv_fprop", C.byref(d), ops.stream())
    _prof_end(e0, "fprop", (m, n_gemm, len(taps) * cin), 2.0 * m[0] * m[1] * m[2] * n_gemm * len(taps) * cin)


def wgrad(g: torch.Tensor, cout: int, x: torch.Tensor, taps: Sequence[Tap], x_stride: int, cin: int,
          m: Tuple[int, int, int], dw: torch.Tensor, ksplit: int = 0, force_bn: int = 0) -> None:
    """dw[co, tap*cin + ci] (fp32) = sum_pixels g[pixel, co] * x[pixel@tap, ci]."""
    _chk(g, "wgrad g")
    _chk(x, "wgrad x")
    assert dw.dtype == F32 and dw.stride(1) == 1
    d = VgConvWgrad()
    d.g, d.g_ld, d.g_coff, d.cout = g.data_ptr(), g.stride(2), 0, cout
    d.x, d.x_n, d.x_h, d.x_w, d.x_ld, d.x_stride = x.data_ptr(), x.shape[0], x.shape[1], x.shape[2], x.stride(2), x_stride
    d.m_n, d.m_h, d.m_w = m
    d.cin, d.num_taps = cin, len(taps)
    _fill_taps(d.taps, taps)
    d.dw, d.dw_ld = dw.data_ptr(), dw.stride(0)
    d.ksplit, d.force_bn = ksplit, force_bn
    e0 = _prof_begin()
    _lib.call("vg_conv_wgrad", C.byref(d), ops.stream())
    _prof_end(e0, "wgrad", (m, cout, len(taps) * cin), 2.0 * m[0] * m[1] * m[2] * cout * len(taps) * cin)


def pad_channels(t: torch.Tensor, c_pad: int) -> torch.Tensor:
    """Widen the logical channel count of an NHWC view to ``c_pad`` (the physical padding must already be there
    and finite -- buffers with ld > c are allocated zeroed)."""
    if t.shape[3] == c_pad:
        return t
    assert t.stride(2) >= c_pad, (t.shape, t.stride(), c_pad)
    return t.as_strided((t.shape[0], t.shape[1], t.shape[2], c_pad), t.stride())


def new_act(n, h, w, c, device, dtype=BF16) -> torch.Tensor:
    """Fresh NHWC activation; channel count padded to a multiple of 64 with zeros when needed."""
    ld = c if c % 64 == 0 else round_up(c, 64)
    if ld == c:
        return torch.empty((n, h, w, c), dtype=dtype, device=device)
    return torch.zeros((n, h, w, ld), dtype=dtype, device=device)[..., :c]


class ConvLinear:
    """A Conv2d-shaped linear map (cin -> cout, kernel kh x kw, stride 1|2, padding) on the tensor pipe.

    Weight matrices are bf16 re-layouts of the fp32 OIHW tensor produced by :meth:`prep_fwd` (used by
    ``forward`` and ``backward_weight`` consumers) and :meth:`prep_bwd` (used by ``backward_data``).
    """

    def __init__(self, cin: int, cout: int, kh: int, kw: int, stride: int = 1, pad: Tuple[int, int] = (0, 0),
                 in_hw: Optional[Tuple[int, int]] = None):
        self.cin, self.cout, self.kh, self.kw, self.s = cin, cout, kh, kw, stride
        self.ph, self.pw = pad
        self.cin_p, self.cout_p = round_up(cin, 64), round_up(cout, 64)
        # "flat": the kernel covers the whole input (output 1x1) -> plain GEMM over the flattened (h, w, c) axis
        self.flat = (in_hw is not None and stride == 1 and pad == (0, 0) and (kh, kw) == tuple(in_hw)
                     and kh * kw > 1 and cin % 64 == 0)
        # "column": kernel (H x 1), output one row -> the data gradient is a pixel shuffle along H
        self.column = (not self.flat and in_hw is not None and stride == 1 and pad == (0, 0) and kw == 1
                       and kh == in_hw[0] and kh > 1)
        self.in_hw = in_hw
        assert stride in (1, 2)

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        return (h + 2 * self.ph - self.kh) // self.s + 1, (w + 2 * self.pw - self.kw) // self.s + 1

    # ---------------------------------------------------------------- weights
    def prep_fwd(self, w: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bf16 [cout, kh*kw*cin_p]  (K order: r, q, ci).  ``scale`` = device scalar sigma -> stores w / sigma."""
        co, ci, kh, kw = w.shape
        out = (torch.zeros if self.cin_p != ci else torch.empty)((co, kh, kw, self.cin_p), dtype=BF16, device=w.device)
        ops.strided_copy(w.permute(0, 2, 3, 1), out[..., :ci], scale, scale_inverse=scale is not None)
        return out.view(co, kh * kw * self.cin_p)

    def prep_bwd(self, w: torch.Tensor, scale: Optional[torch.Tensor] = None) -> Dict:
        """Weight matrices of the data gradient (rows = cin of the conv, K = taps x cout_p)."""
        co, ci, kh, kw = w.shape
        inv = scale is not None
        mk = torch.zeros if self.cout_p != co else torch.empty
        if self.flat or self.column or (self.s == 2 and (kh, kw) == (2, 2) and (self.ph, self.pw) == (0, 0)):
            # single tap, GEMM N = (r, q, ci): [(r,q,ci)][co_p]
            out = mk((kh, kw, ci, self.cout_p), dtype=BF16, device=w.device)
            ops.strided_copy(w.permute(2, 3, 1, 0), out[..., :co], scale, scale_inverse=inv)
            return {"shuffle": out.view(kh * kw * ci, self.cout_p)}
        if self.s == 1:
            out = mk((ci, kh, kw, self.cout_p), dtype=BF16, device=w.device)
            ops.strided_copy(w.permute(1, 2, 3, 0), out[..., :co], scale, scale_inverse=inv)
            return {"s1": out.view(ci, kh * kw * self.cout_p)}
        mats = {}
        for a in (0, 1):
            for b in (0, 1):
                r0, q0 = (a + self.ph) % 2, (b + self.pw) % 2
                sub = w[:, :, r0::2, q0::2]
                nr, nq = sub.shape[2], sub.shape[3]
                out = mk((ci, nr, nq, self.cout_p), dtype=BF16, device=w.device)
                ops.strided_copy(sub.permute(1, 2, 3, 0), out[..., :co], scale, scale_inverse=inv)
                taps = [(0, (b + self.pw - q0) // 2 - kq, 0, (a + self.ph - r0) // 2 - kr)
                        for kr in range(nr) for kq in range(nq)]
                mats[(a, b)] = (out.view(ci, nr * nq * self.cout_p), taps)
        return {"parity": mats}

    # ---------------------------------------------------------------- primitives
    def forward(self, x: torch.Tensor, wf: torch.Tensor, bias=None, act: int = 0, out: Optional[torch.Tensor] = None,
                out_kind: int = 0) -> torch.Tensor:
        n, h, w, _ = x.shape
        oh, ow = self.out_hw(h, w)
        if out is None:
            out = (new_act(n, oh, ow, self.cout, x.device) if out_kind == 0
                   else torch.zeros((n, oh, ow, self.cout), dtype=F32, device=x.device))
        if self.flat:
            assert x.is_contiguous()
            xf = x.view(n, 1, 1, h * w * self.cin)
            fprop(xf, [(0, 0, 0, 0)], 1, h * w * self.cin, wf, self.cout, (n, 1, 1), out, out_kind=out_kind, bias=bias,
                  act=act)
            return out
        xp = pad_channels(x, self.cin_p)
        taps = conv_taps(self.kh, self.kw, self.s, self.ph, self.pw, xp.stride(2))
        fprop(xp, taps, self.s, self.cin_p, wf, self.cout, (n, oh, ow), out, out_kind=out_kind, bias=bias, act=act)
        return out

    def backward_data(self, dy: torch.Tensor, wb: Dict, in_hw: Tuple[int, int], bias=None, act: int = 0,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n, oh, ow, _ = dy.shape
        h, w = in_hw
        if out is None:
            out = new_act(n, h, w, self.cin, dy.device)
        g = pad_channels(dy, self.cout_p)
        if "shuffle" in wb:
            # pixel shuffle: GEMM column (r, q, ci) of input pixel (oh, ow) lands at output pixel (oh*kh + r, ow*kw + q).
            # Covers the full-kernel case (1x1 input), the column kernel (kh x 1) and 2x2 stride 2.
            fprop(g, [(0, 0, 0, 0)], 1, self.cout_p, wb["shuffle"], self.kh * self.kw * self.cin, (n, oh, ow), out,
                  su=(self.kh, self.kw), cout_per_sub=self.cin, bias=bias, act=act)
            return out
        if "s1" in wb:
            taps = [(0, self.pw - q, 0, self.ph - r) for r in range(self.kh) for q in range(self.kw)]
            fprop(g, taps, 1, self.cout_p, wb["s1"], self.cin, (n, h, w), out, bias=bias, act=act)
            return out
        for (a, b), (mat, taps) in wb["parity"].items():
            fprop(g, taps, 1, self.cout_p, mat, self.cin, (n, h // 2, w // 2), out, su=(2, 2), sub0=(a, b),
                  cout_per_sub=self.cin, bias=bias, act=act)
        return out

    def backward_weight(self, dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """fp32 gradient in OIHW order as a (permuted) view [cout, cin, kh, kw] of the kernel's [cout][r][q][ci] result."""
        n, oh, ow, _ = dy.shape
        _, h, w, _ = x.shape
        if self.flat:
            assert x.is_contiguous()
            k = h * w * self.cin
            dw = torch.empty((self.cout, k), dtype=F32, device=x.device)
            wgrad(dy, self.cout, x.view(n, 1, 1, k), [(0, 0, 0, 0)], 1, k, (n, 1, 1), dw)
            return dw.view(self.cout, self.kh, self.kw, self.cin).permute(0, 3, 1, 2)
        xp = pad_channels(x, self.cin_p)
        taps = conv_taps(self.kh, self.kw, self.s, self.ph, self.pw, xp.stride(2))
        dw = torch.empty((self.cout, len(taps) * self.cin_p), dtype=F32, device=x.device)
        wgrad(dy, self.cout, xp, taps, self.s, self.cin_p, (n, oh, ow), dw)
        return dw.view(self.cout, self.kh, self.kw, self.cin_p)[..., :self.cin].permute(0, 3, 1, 2)
