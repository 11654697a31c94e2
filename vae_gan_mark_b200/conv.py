"""Geometry of the tensor-core implicit GEMMs: turns conv / conv-transpose layer shapes into the tap
tables and descriptors of ``vg_conv_fprop`` / ``vg_conv_wgrad`` (include/vaegan_b200.h).

Activations are NHWC bf16 torch tensors wrapped in :class:`Act` (a view into a possibly wider
buffer: ``buf[..., coff:coff+c]``), so a producer can write straight into a channel slice of the concat
buffer of a U-Net skip (reference: torch.cat at vae-gan-v2.py:251-274).

Every ConvTranspose2d is handled as the adjoint of the Conv2d with the same weight tensor: its forward
is that conv's data-gradient, its data-gradient is that conv's forward, and its weight gradient is that
conv's weight gradient with the two operands swapped.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import VgConvFprop, VgConvWgrad

BF16 = torch.bfloat16


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


@dataclass
class Act:
    """NHWC bf16 activation: channels [coff, coff+c) of ``buf`` (shape [N, H, W, ld])."""
    buf: torch.Tensor
    c: int
    coff: int = 0

    @staticmethod
    def empty(n, h, w, c, ld: Optional[int] = None, device="cuda", zero=False) -> "Act":
        ld = ld or round_up(c, 8)
        mk = torch.zeros if (zero or ld != c) else torch.empty
        return Act(mk((n, h, w, ld), dtype=BF16, device=device), c, 0)

    @property
    def n(self): return self.buf.shape[0]
    @property
    def h(self): return self.buf.shape[1]
    @property
    def w(self): return self.buf.shape[2]
    @property
    def ld(self): return self.buf.shape[3]
    @property
    def ptr(self): return self.buf.data_ptr()

    def slice(self, coff: int, c: int) -> "Act":
        return Act(self.buf, c, self.coff + coff)

    def view(self) -> torch.Tensor:
        """Strided torch view [N,H,W,c] (for tests / glue only)."""
        return self.buf[..., self.coff:self.coff + self.c]

    def reshape(self, n, h, w, c) -> "Act":
        """Reinterpret a dense (ld == c, coff == 0) activation with new NHWC dims of equal size."""
        assert self.coff == 0 and self.ld == self.c and n * h * w * c == self.buf.numel()
        return Act(self.buf.view(n, h, w, c), c, 0)


Tap = Tuple[int, int, int, int]   # (c_base, dw, sh, dh)


def conv_taps(kh: int, kw: int, stride: int, ph: int, pw: int, ld: int, coff: int) -> List[Tap]:
    """Taps of a (kh x kw, stride, pad) convolution reading its input through the stride view."""
    taps = []
    for r in range(kh):
        for q in range(kw):
            rr, qq = r - ph, q - pw
            if stride == 1:
                taps.append((coff, qq, 0, rr))
            else:
                taps.append(((qq % 2) * ld + coff, qq // 2, rr % 2, rr // 2))
    return taps


def _fill_taps(dst, taps: Sequence[Tap]):
    assert len(taps) <= _lib.VG_MAX_TAPS, f"{len(taps)} taps > {_lib.VG_MAX_TAPS}"
    for i, t in enumerate(taps):
        for j in range(4):
            dst[i][j] = int(t[j])


def fprop(x: Act, taps: Sequence[Tap], x_stride: int, cin: int, w: torch.Tensor, n_gemm: int, m: Tuple[int, int, int],
          out: torch.Tensor, out_hw: Tuple[int, int], out_ld: int, out_coff: int = 0, out_kind: int = 0,
          su: Tuple[int, int] = (1, 1), sub0: Tuple[int, int] = (0, 0), cout_per_sub: Optional[int] = None,
          bias: Optional[torch.Tensor] = None, act: int = 0, ksplit: int = 0, force_bn: int = 0) -> None:
    """Launch vg_conv_fprop.  ``w`` is bf16 [n_gemm, len(taps)*cin] (row stride w.stride(0))."""
    assert w.dtype == BF16 and w.stride(1) == 1 and x.buf.dtype == BF16 and x.buf.is_contiguous()
    d = VgConvFprop()
    d.x, d.x_n, d.x_h, d.x_w, d.x_ld, d.x_stride = x.ptr, x.n, x.h, x.w, x.ld, x_stride
    d.m_n, d.m_h, d.m_w = m
    d.cin, d.num_taps = cin, len(taps)
    _fill_taps(d.taps, taps)
    d.w, d.w_ld, d.n_gemm = w.data_ptr(), w.stride(0), n_gemm
    d.out, d.out_kind = out.data_ptr(), out_kind
    d.out_h, d.out_w, d.out_ld, d.out_coff = out_hw[0], out_hw[1], out_ld, out_coff
    d.su_h, d.su_w = su
    d.sub_h0, d.sub_w0 = sub0
    d.cout_per_sub = cout_per_sub or n_gemm
    d.bias = bias.data_ptr() if bias is not None else None
    d.act, d.ksplit, d.force_bn = act, ksplit, force_bn
    _lib.call("vg_conv_fprop", C.byref(d), _stream())


def wgrad(g: Act, cout: int, x: Act, taps: Sequence[Tap], x_stride: int, cin: int, m: Tuple[int, int, int],
          dw: torch.Tensor, ksplit: int = 0, force_bn: int = 0) -> None:
    """Launch vg_conv_wgrad: dw[cout, len(taps)*cin] (fp32) = sum_pixels g^T x@tap."""
    assert dw.dtype == torch.float32 and dw.stride(1) == 1
    d = VgConvWgrad()
    d.g, d.g_ld, d.g_coff, d.cout = g.ptr, g.ld, g.coff, cout
    d.x, d.x_n, d.x_h, d.x_w, d.x_ld, d.x_stride = x.ptr, x.n, x.h, x.w, x.ld, x_stride
    d.m_n, d.m_h, d.m_w = m
    d.cin, d.num_taps = cin, len(taps)
    _fill_taps(d.taps, taps)
    d.dw, d.dw_ld = dw.data_ptr(), dw.stride(0)
    d.ksplit, d.force_bn = ksplit, force_bn
    _lib.call("vg_conv_wgrad", C.byref(d), _stream())
