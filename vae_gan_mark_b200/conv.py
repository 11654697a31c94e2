"""Tensor-core convolutions: geometry -> tap tables -> ``vg_conv_fprop`` / ``vg_conv_wgrad`` launches.

Activations are NHWC torch views ([N, H, W, C], last stride 1, pixel stride ``ld = stride(2)`` >= C so a tensor
may be a channel slice of a wider concat buffer), bf16 by default.  ``ConvLinear`` describes one Conv2d-shaped
linear map and provides its three primitives (forward, data gradient, weight gradient); a ConvTranspose2d is the
adjoint of the Conv2d with the same weight tensor, so its forward is that conv's data gradient, its data gradient
is that conv's forward and its weight gradient is that conv's weight gradient with the operands swapped (IOHW of
the transpose == OIHW of its adjoint).  Reference layers: vae-gan.py:52-60,76-81,153-157;
vae-gan-v2.py:123-127,168-176,199-241; vae-gan-unet.py:148-154,194-221.

High-accuracy mode (fp32 activations, ``ops.set_precision("fp32")``): the tensor core only takes bf16, so each fp32
operand is split into three bf16 planes hi + mid + lo (``vg_split3``) and the six plane pairs with the largest
products (hi.hi, hi.mid, mid.hi, mid.mid, hi.lo, lo.hi) are accumulated in the fp32 TMEM accumulator -- the K loop
simply runs over (pair, tap) "virtual taps".  The result carries ~2^-22 relative error per product instead of 2^-8.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from ._lib import VgConvFprop, VgConvWgrad
from .dispatch import launch_op
from .ops import BF16, F32, round_up

Tap = Tuple[int, int, int, int]   # (c_base, dw, sh, dh)
PAIRS = ((0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0))   # (activation plane, weight plane)

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


def _chk(t: torch.Tensor, what: str):
    assert t.dtype == BF16 and ops.nhwc_ok(t), f"{what}: not an NHWC bf16 view {tuple(t.shape)} {t.stride()}"


import os as _os
_PREFER_MN_MODE = _os.environ.get("VG_PREFER_MN", "")
HALO_MODE = 0      # VgConvFprop.halo_mode: 0 auto, -1 never, 1 force (tests)
# Deterministic split reductions (bf16 mode): the weight-gradient kernel stores each pixel split's partial tile into a
# scratch buffer and adds them in a fixed order (two-stage reduce) instead of fp32 atomics, and split-K forward launches
# (the 2z-wide heads GEMM) run unsplit.  Results are then bit-identical from run to run; costs a few percent.
DETERMINISTIC = False


def fprop(x: torch.Tensor, taps: Sequence[Tap], x_stride: int, cin: int, w: torch.Tensor, n_gemm: int,
          m: Tuple[int, int, int], out: torch.Tensor, out_kind: int = 0, su: Tuple[int, int] = (1, 1),
          sub0: Tuple[int, int] = (0, 0), cout_per_sub: Optional[int] = None, bias: Optional[torch.Tensor] = None,
          act: int = 0, ksplit: int = 0, force_bn: int = 0, wk: Optional[Sequence[int]] = None,
          b_mn_major: bool = False, groups: Optional[Sequence[Tuple[int, Tuple[int, int]]]] = None,
          halo_mode: Optional[int] = None, stats: Optional[torch.Tensor] = None) -> None:
    """out[pixel, n] = sum_{tap, c<cin} x[pixel@tap, c] * w[n, wk[tap] + c]  (wk[tap] = tap*cin by default).
    ``out`` is an NHWC view ([N, OH, OW, C']); ``x`` and ``w`` are bf16.
    ``b_mn_major``: w is [K rows (c), columns] and out[pixel, n] = sum x[pixel@tap, c] * w[c, wk[tap] + n].
    ``groups``: [(number of taps, (sub_h0, sub_w0)), ...] -- up to 4 problems in one launch, taps listed group by group.
    ``stats``: fp32 [1, 2, channels] that receives the per-channel sum / sum of squares of the stored output (the batch
    statistics of a following BatchNorm2d), computed in the epilogue.
    The launch itself is the dispatcher op ``torch.ops.vaegan.vg_conv_fprop`` (tap / group tables flattened to int[])."""
    _chk(x, "fprop x")
    assert w.dtype == BF16 and w.stride(1) == 1 and out.stride(3) == 1
    assert len(taps) <= _lib.VG_MAX_FPROP_TAPS, f"{len(taps)} taps > {_lib.VG_MAX_FPROP_TAPS}"
    assert wk is not None or w.shape[1] == len(taps) * cin, (w.shape, len(taps), cin)
    assert not b_mn_major or wk is not None
    if DETERMINISTIC and out_kind == 2 and ksplit == 0 and x.dtype == BF16:
        ksplit = 1      # one split adds into the zeroed destination exactly once: order-independent
    flat_groups = None
    if groups is not None:
        assert 2 <= len(groups) <= 4 and sum(g[0] for g in groups) == len(taps) and ksplit in (0, 1)
        flat_groups = [int(v) for nt, (sh, sw) in groups for v in (nt, sh, sw)]
    e0 = _prof_begin()
    _fprop_launch(x, [int(v) for t in taps for v in t], x_stride, cin, w, n_gemm, list(m), out, out_kind, list(su), list(sub0),
                  cout_per_sub or n_gemm, bias, act, ksplit, force_bn, [int(v) for v in wk] if wk is not None else None,
                  b_mn_major, flat_groups, HALO_MODE if halo_mode is None else halo_mode, stats)
    _prof_end(e0, "fprop", (m, n_gemm, len(taps) * cin), 2.0 * m[0] * m[1] * m[2] * n_gemm * len(taps) * cin)


@launch_op("vg_conv_fprop(Tensor x, int[] taps, int x_stride, int cin, Tensor w, int n_gemm, int[] m, Tensor(a!) out, int out_kind, "
           "int[] su, int[] sub0, int cout_per_sub, Tensor? bias, int act, int ksplit, int force_bn, int[]? wk, bool b_mn_major, "
           "int[]? groups, int halo_mode, Tensor(b!)? stats) -> ()")
def _fprop_launch(x, taps, x_stride, cin, w, n_gemm, m, out, out_kind, su, sub0, cout_per_sub, bias, act, ksplit, force_bn, wk,
                  b_mn_major, groups, halo_mode, stats):
    """Descriptor marshalling + launch of the forward / data-gradient tensor-core kernel (include/vaegan_b200.h: VgConvFprop)."""
    ntaps = len(taps) // 4
    d = VgConvFprop()
    d.x, d.x_n, d.x_h, d.x_w, d.x_ld, d.x_stride = x.data_ptr(), x.shape[0], x.shape[1], x.shape[2], x.stride(2), x_stride
    d.m_n, d.m_h, d.m_w = m
    d.cin, d.num_taps = cin, ntaps
    for i in range(ntaps):
        for j in range(4):
            d.taps[i][j] = taps[4 * i + j]
    if wk is not None:
        d.use_wk = 1
        for i, v in enumerate(wk):
            d.wk[i] = v
    d.w, d.w_ld, d.n_gemm = w.data_ptr(), w.stride(0), n_gemm
    d.out, d.out_kind = out.data_ptr(), out_kind
    d.out_h, d.out_w, d.out_ld, d.out_coff = out.shape[1], out.shape[2], out.stride(2), 0
    d.su_h, d.su_w = su
    d.sub_h0, d.sub_w0 = sub0
    d.cout_per_sub = cout_per_sub
    d.bias = bias.data_ptr() if bias is not None else None
    d.act, d.ksplit, d.force_bn = act, ksplit, force_bn
    d.b_mn_major, d.w_rows = int(b_mn_major), w.shape[0]
    d.halo_mode = halo_mode
    if stats is not None:
        assert stats.dtype == F32 and stats.is_contiguous() and stats.numel() == 2 * d.cout_per_sub and out_kind != 2
        d.stats = stats.data_ptr()
    if groups is not None:
        d.num_groups = len(groups) // 3
        for i in range(d.num_groups):
            d.group_ntaps[i] = groups[3 * i]
            d.group_sub[i][0], d.group_sub[i][1] = groups[3 * i + 1], groups[3 * i + 2]
    _lib.call("vg_conv_fprop", C.byref(d), ops.stream())


def wgrad(g: torch.Tensor, cout: int, x: torch.Tensor, taps: Sequence[Tap], x_stride: int, cin: int,
          m: Tuple[int, int, int], dw: torch.Tensor, ksplit: int = 0, force_bn: int = 0,
          pairs: Optional[Sequence[Tuple[int, int]]] = None) -> None:
    """dw[co, tap*cin + ci] (fp32) = sum_pixels g[pixel, co] * x[pixel@tap, ci]; ``pairs`` = channel offsets (g, x) of
    split-precision operand planes accumulated into the same dw.  The launch is ``torch.ops.vaegan.vg_conv_wgrad``."""
    _chk(g, "wgrad g")
    _chk(x, "wgrad x")
    assert dw.dtype == F32 and dw.stride(1) == 1 and len(taps) <= _lib.VG_MAX_TAPS
    e0 = _prof_begin()
    _wgrad_launch(g, cout, x, [int(v) for t in taps for v in t], x_stride, cin, list(m), dw, ksplit, force_bn,
                  [int(v) for ab in pairs for v in ab] if pairs else None, DETERMINISTIC)
    _prof_end(e0, "wgrad", (m, cout, len(taps) * cin),
              2.0 * m[0] * m[1] * m[2] * cout * len(taps) * cin * (len(pairs) if pairs else 1))


@launch_op("vg_conv_wgrad(Tensor g, int cout, Tensor x, int[] taps, int x_stride, int cin, int[] m, Tensor(a!) dw, int ksplit, "
           "int force_bn, int[]? pairs, bool deterministic) -> ()")
def _wgrad_launch(g, cout, x, taps, x_stride, cin, m, dw, ksplit, force_bn, pairs, deterministic):
    """Descriptor marshalling + launch of the weight-gradient tensor-core kernel (include/vaegan_b200.h: VgConvWgrad)."""
    ntaps = len(taps) // 4
    d = VgConvWgrad()
    d.g, d.g_ld, d.g_coff, d.cout = g.data_ptr(), g.stride(2), 0, cout
    d.x, d.x_n, d.x_h, d.x_w, d.x_ld, d.x_stride = x.data_ptr(), x.shape[0], x.shape[1], x.shape[2], x.stride(2), x_stride
    d.m_n, d.m_h, d.m_w = m
    d.cin, d.num_taps = cin, ntaps
    for i in range(ntaps):
        for j in range(4):
            d.taps[i][j] = taps[4 * i + j]
    if pairs:
        d.num_combos = len(pairs) // 2
        for i in range(d.num_combos):
            d.combo_g[i], d.combo_x[i] = pairs[2 * i], pairs[2 * i + 1]
    d.dw, d.dw_ld = dw.data_ptr(), dw.stride(0)
    d.ksplit, d.force_bn = ksplit, force_bn
    if deterministic:
        need = int(_lib.lib().vg_conv_wgrad_workspace(C.byref(d)))
        assert need >= 0, "vg_conv_wgrad_workspace rejected the descriptor"
        if need > 0:
            ws = torch.empty(need // 4, dtype=F32, device=dw.device)      # lives until the stream has consumed it (caching allocator)
            d.workspace, d.workspace_bytes = ws.data_ptr(), need
    _lib.call("vg_conv_wgrad", C.byref(d), ops.stream())


def pad_channels(t: torch.Tensor, c_pad: int) -> torch.Tensor:
    """Widen the logical channel count of an NHWC view to ``c_pad`` (the physical padding must already be there
    and finite -- buffers with ld > c are allocated zeroed)."""
    if t.shape[3] == c_pad:
        return t
    assert t.stride(2) >= c_pad, (t.shape, t.stride(), c_pad)
    return t.as_strided((t.shape[0], t.shape[1], t.shape[2], c_pad), t.stride())


def new_act(n, h, w, c, device, dtype=None) -> torch.Tensor:
    """Fresh NHWC activation; channel count padded to a multiple of 64 with zeros when needed."""
    dtype = dtype or ops.act_dtype()
    ld = c if c % 64 == 0 else round_up(c, 64)
    if ld == c:
        return torch.empty((n, h, w, c), dtype=dtype, device=device)
    return torch.zeros((n, h, w, ld), dtype=dtype, device=device)[..., :c]


def _operand(view: torch.Tensor, kp: int, hi: bool, scale=None) -> torch.Tensor:
    """Weight operand from a [rows, <tap dims...>, k]-shaped (permuted, possibly strided) fp32 view of the master
    weight.  bf16 mode: bf16 [rows, taps*kp]; high-accuracy mode: bf16 [rows, taps*3*kp], planes innermost per tap."""
    rows, k = view.shape[0], view.shape[-1]
    taps = 1
    for t in view.shape[1:-1]:
        taps *= t
    inv = scale is not None
    if not hi:
        out = (torch.zeros if kp != k else torch.empty)(tuple(view.shape[:-1]) + (kp,), dtype=BF16, device=view.device)
        ops.strided_copy(view, out[..., :k], scale, scale_inverse=inv)
        return out.view(rows, taps * kp)
    tmp = torch.empty(tuple(view.shape), dtype=F32, device=view.device)
    ops.strided_copy(view, tmp, scale, scale_inverse=inv)
    return ops.split3(tmp.view(1, rows, taps, k)).view(rows, taps * 3 * kp)


HI_STEPS_PER_SPLIT = 16   # high-accuracy mode: K steps (of 64) accumulated in TMEM before an fp32 (RN) add in L2


# K steps one TMEM accumulation may run in a high-accuracy launch; None = HI_STEPS_PER_SPLIT.  The GRU's time-parallel GEMMs
# (layers.GRULayerFn) raise it in bf16 mode: their ~1e-5 truncation bias is far below the bf16 noise of everything around
# them, and an unsplit launch stores its tile directly instead of combining partial tiles with atomics.
HI_MAX_STEPS = None


def _hi_launch(x, vt, stride, cin, w, n_gemm, m, out, wk, bias, act, **kw):
    """High-accuracy launch: the tensor core's fp32 accumulator truncates on every MMA, which biases long K sums by
    ~1e-4; so K is cut into short splits whose partial sums are combined with round-to-nearest fp32 atomics.  A launch
    short enough for ONE accumulation stores its tile directly (bias / activation fused, no zero-fill, no atomics)."""
    ksteps = len(vt) * (cin // 64)
    ksplit = max(1, -(-ksteps // (HI_MAX_STEPS or HI_STEPS_PER_SPLIT)))
    if ksplit == 1 and out.dtype == F32 and not kw.get("groups"):
        fprop(x, vt, stride, cin, w, n_gemm, m, out, out_kind=1, bias=bias, act=act, wk=wk, ksplit=1, **kw)
        return
    out.zero_()
    fprop(x, vt, stride, cin, w, n_gemm, m, out, out_kind=2, bias=bias, act=0, wk=wk, ksplit=ksplit, **kw)
    if act:
        ops.act_fwd_(out, act)


def _vtaps(taps: Sequence[Tap], kp_act: int, kp_w: int, hi: bool):
    """(virtual taps, weight column offsets) of the K loop."""
    if not hi:
        return list(taps), None
    vt, wk = [], []
    for (pa, pb) in PAIRS:
        for i, (cb, dw, sh, dh) in enumerate(taps):
            vt.append((cb + pa * kp_act, dw, sh, dh))
            wk.append((i * 3 + pb) * kp_w)
    return vt, wk


def _hi_wgrad_split(n: int, oh: int, ow: int) -> int:
    """Split factor of a high-accuracy wgrad launch: <= HI_STEPS_PER_SPLIT K steps (64 pixels x one plane pair) each."""
    w = 1
    while w < ow and w < 64:
        w <<= 1
    h = 1
    while h < oh and w * h < 64:
        h <<= 1
    tn = 64 // (w * h)
    tiles = -(-n // tn) * -(-oh // h) * -(-ow // w) * len(PAIRS)
    return max(1, -(-tiles // HI_STEPS_PER_SPLIT))


class ConvLinear:
    """A Conv2d-shaped linear map (cin -> cout, kernel kh x kw, stride 1|2, padding) on the tensor pipe.

    Weight operands are bf16 re-layouts of the fp32 OIHW tensor produced by :meth:`prep_fwd` (used by ``forward``)
    and :meth:`prep_bwd` (used by ``backward_data``); ``hi`` selects the split-precision layout.
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
        self.shuffle = self.flat or self.column or (stride == 2 and (kh, kw) == (2, 2) and pad == (0, 0))
        self.in_hw = in_hw
        assert stride in (1, 2)

    def prefer_mn(self, gemm_pixels: int) -> bool:
        """Data gradient in bf16 mode: read the forward operand as an MN-major B operand (no transposed weight copy per
        step) or make the transposed K-major copy first?  With the warp-uniform issue loops and CTA pairs on both paths
        the MN-major operand costs nothing up to 64x64x64 pixels (tools/gpu_mn_vs_k.py: 1618 vs 1610 TFLOP/s on the
        512 -> 512 3x3 layer) and ~7 % on the largest launches (1510 vs 1624 at 128x128x64: two boxes per CTA and K step
        instead of one); the copy costs ~5 us of launch plus the weight at ~0.8 TB/s.  K-major also keeps the halo mode
        of the narrow 3x3 layers (data gradient with <= 64 input channels), which the MN-major path does not have."""
        if self.shuffle and self.cin != self.cin_p:
            # the pixel-shuffle GEMM's columns are (r, q, ci) with ci < cin, but the forward operand's columns are
            # (r, q, ci_p) zero-padded to 64 per tap: no uniform column map exists (up_tconv3 64 -> 32 of vae-gan-oldv.py)
            return False
        if self.s == 1 and (self.kh, self.kw) == (3, 3) and self.cin <= 64 and gemm_pixels >= 65536:
            return False                                   # halo mode (K-major weights only): 892 vs 671 TFLOP/s
        nk = self.cin * self.cout * self.kh * self.kw
        if _PREFER_MN_MODE == "old":       # development A/B: the mid-round cost model (MN-major assumed 40 % slower)
            return 0.4 * gemm_pixels * nk / 1.4e9 < 5.0 * (4 if (self.s == 2 and not self.shuffle) else 1) + nk * 6 / 1e6
        gemm_us = 2.0 * gemm_pixels * nk / 1.5e9
        slowdown_us = 0.07 * gemm_us if gemm_pixels >= (1 << 19) else 0.0
        copies = 4 if (self.s == 2 and not self.shuffle) else 1
        copy_us = 5.0 * copies + nk * 6 / 0.8e6
        return slowdown_us < copy_us

    def _parity_taps(self, col_step: int):
        """Taps, weight-column offsets and groups of the stride-2 data gradient: output pixel (2i + a, 2j + b) collects
        the kernel taps (r, q) with r = a + ph, q = b + pw (mod 2), reading dy at (i + (a + ph - r) / 2, ...)."""
        taps, wk, groups = [], [], []
        for a in (0, 1):
            for b in (0, 1):
                r0, q0 = (a + self.ph) % 2, (b + self.pw) % 2
                rs, qs = range(r0, self.kh, 2), range(q0, self.kw, 2)
                taps += [(0, (b + self.pw - q) // 2, 0, (a + self.ph - r) // 2) for r in rs for q in qs]
                wk += [(r * self.kw + q) * col_step for r in rs for q in qs]
                groups.append((len(rs) * len(qs), (a, b)))
        assert all(nt > 0 for nt, _ in groups), "stride-2 data gradient needs a kernel of at least 2x2"
        return taps, wk, groups

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        return (h + 2 * self.ph - self.kh) // self.s + 1, (w + 2 * self.pw - self.kw) // self.s + 1

    # ---------------------------------------------------------------- weights
    def prep_fwd(self, w: torch.Tensor, scale: Optional[torch.Tensor] = None, hi: bool = False) -> torch.Tensor:
        """rows = cout, K order (r, q, ci).  ``scale`` = device scalar sigma -> stores w / sigma."""
        co, ci, kh, kw = w.shape
        if self.flat and hi:
            # one K axis of length kh*kw*ci: lay the fp32 weight out as [co][(r, q, ci)] first, then split it
            tmp = torch.empty((co, kh, kw, ci), dtype=F32, device=w.device)
            ops.strided_copy(w.permute(0, 2, 3, 1), tmp, scale, scale_inverse=scale is not None)
            return _operand(tmp.view(co, 1, kh * kw * ci), kh * kw * ci, True)
        return _operand(w.permute(0, 2, 3, 1), self.cin_p, hi, scale)

    def prep_bwd(self, w: torch.Tensor, scale: Optional[torch.Tensor] = None, hi: bool = False) -> Dict:
        """Weight operands of the data gradient (rows = cin of the conv side, K = taps x cout_p)."""
        co, ci, kh, kw = w.shape
        if self.shuffle:
            # single tap, GEMM N = (r, q, ci): rows (r, q, ci), K = co
            # (the leading dims of the permuted view are only a factorisation of the row index (r, q, ci))
            return {"shuffle": _operand(w.permute(2, 3, 1, 0), self.cout_p, hi, scale).view(kh * kw * ci, -1)}
        if self.s == 1:
            return {"s1": _operand(w.permute(1, 2, 3, 0), self.cout_p, hi, scale)}
        if not hi and all(len(range(r0, self.kh, 2)) > 0 for r0 in (0, 1)) and all(len(range(q0, self.kw, 2)) > 0 for q0 in (0, 1)):
            return {"parity_k": _operand(w.permute(1, 2, 3, 0), self.cout_p, False, scale)}
        mats = {}
        for a in (0, 1):
            for b in (0, 1):
                r0, q0 = (a + self.ph) % 2, (b + self.pw) % 2
                sub = w[:, :, r0::2, q0::2]
                nr, nq = sub.shape[2], sub.shape[3]
                taps = [(0, (b + self.pw - q0) // 2 - kq, 0, (a + self.ph - r0) // 2 - kr)
                        for kr in range(nr) for kq in range(nq)]
                mats[(a, b)] = (_operand(sub.permute(1, 2, 3, 0), self.cout_p, hi, scale), taps)
        return {"parity": mats}

    # ---------------------------------------------------------------- primitives
    def forward(self, x: torch.Tensor, wf: torch.Tensor, bias=None, act: int = 0, out: Optional[torch.Tensor] = None,
                out_kind: Optional[int] = None, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``stats`` (bf16 mode only): see :func:`fprop`."""
        n, h, w, _ = x.shape
        hi = x.dtype == F32
        assert stats is None or (not hi and not self.flat)
        oh, ow = self.out_hw(h, w)
        if out_kind is None:
            out_kind = 1 if hi else 0
        if out is None:
            out = (new_act(n, oh, ow, self.cout, x.device, x.dtype) if out_kind != 2
                   else torch.zeros((n, oh, ow, self.cout), dtype=F32, device=x.device))
        if self.flat:
            assert x.is_contiguous()
            k = h * w * self.cin
            xf = x.view(n, 1, 1, k)
            xs = ops.split3(xf) if hi else xf
            vt, wk = _vtaps([(0, 0, 0, 0)], k, k, hi)
            if hi:
                _hi_launch(xs, vt, 1, k, wf, self.cout, (n, 1, 1), out, wk, bias, act)
            elif out_kind == 0 and bias is None and act == 0 and k >= 64 * 256:
                # a handful of output tiles against a very long K axis (the data gradient of the bottleneck
                # ConvTranspose2d, vae-gan-unet.py:194: K = 16*16*1024): without split-K three to five CTAs would stream
                # the whole weight (1.26 ms at 256x256); accumulate the K splits in fp32 and round once
                acc = torch.zeros((n, 1, 1, self.cout), dtype=F32, device=x.device)
                fprop(xs, vt, 1, k, wf, self.cout, (n, 1, 1), acc, out_kind=2, wk=wk)
                ops.strided_copy(acc, out)
            else:
                fprop(xs, vt, 1, k, wf, self.cout, (n, 1, 1), out, out_kind=out_kind, bias=bias, act=act, wk=wk)
            return out
        xs = ops.split3(x) if hi else pad_channels(x, self.cin_p)
        taps = conv_taps(self.kh, self.kw, self.s, self.ph, self.pw, xs.stride(2))
        vt, wk = _vtaps(taps, self.cin_p, self.cin_p, hi)
        if hi:
            _hi_launch(xs, vt, self.s, self.cin_p, wf, self.cout, (n, oh, ow), out, wk, bias, act)
        else:
            fprop(xs, vt, self.s, self.cin_p, wf, self.cout, (n, oh, ow), out, out_kind=out_kind, bias=bias, act=act, wk=wk,
                  stats=stats)
        return out

    def backward_data(self, dy: torch.Tensor, wb: Dict, in_hw: Tuple[int, int], bias=None, act: int = 0,
                      out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``stats`` (bf16 mode, single-launch variants only): see :func:`fprop` -- used when this primitive is the
        FORWARD of a ConvTranspose2d that feeds a BatchNorm2d."""
        n, oh, ow, _ = dy.shape
        h, w = in_hw
        hi = dy.dtype == F32
        assert stats is None or (not hi and "parity" not in wb)
        kind = 1 if hi else 0
        if out is None:
            out = new_act(n, h, w, self.cin, dy.device, dy.dtype)
        g = ops.split3(dy) if hi else pad_channels(dy, self.cout_p)
        if "mn" in wb:
            # bf16 mode: the forward operand [cout][(r, q, ci_p)] is read as the MN-major B operand (K = cout rows,
            # N = ci columns of one tap); no transposed copy of the weights exists
            assert not hi
            wf = wb["mn"]
            if self.shuffle:
                fprop(g, [(0, 0, 0, 0)], 1, self.cout_p, wf, self.kh * self.kw * self.cin, (n, oh, ow), out, out_kind=kind,
                      su=(self.kh, self.kw), cout_per_sub=self.cin, bias=bias, act=act, wk=[0], b_mn_major=True, stats=stats)
            elif self.s == 1:
                taps = [(0, self.pw - q, 0, self.ph - r) for r in range(self.kh) for q in range(self.kw)]
                wk = [(r * self.kw + q) * self.cin_p for r in range(self.kh) for q in range(self.kw)]
                fprop(g, taps, 1, self.cout_p, wf, self.cin, (n, h, w), out, out_kind=kind, bias=bias, act=act, wk=wk,
                      b_mn_major=True, stats=stats)
            else:
                taps, wk, groups = self._parity_taps(self.cin_p)
                fprop(g, taps, 1, self.cout_p, wf, self.cin, (n, h // 2, w // 2), out, out_kind=kind, su=(2, 2),
                      cout_per_sub=self.cin, bias=bias, act=act, wk=wk, b_mn_major=True, groups=groups, stats=stats)
            return out
        if "parity_k" in wb:
            # K-major operand [cin][(r, q)][cout_p] holding every tap; the four output-parity classes are the four
            # groups of ONE launch, each reading its own taps through the column offsets
            taps, wk, groups = self._parity_taps(self.cout_p)
            fprop(g, taps, 1, self.cout_p, wb["parity_k"], self.cin, (n, h // 2, w // 2), out, out_kind=kind, su=(2, 2),
                  cout_per_sub=self.cin, bias=bias, act=act, wk=wk, groups=groups, stats=stats)
            return out
        if "shuffle" in wb:
            # pixel shuffle: GEMM column (r, q, ci) of input pixel (oh, ow) lands at output pixel (oh*kh + r, ow*kw + q).
            # Covers the full-kernel case (1x1 input), the column kernel (kh x 1) and 2x2 stride 2.
            vt, wk = _vtaps([(0, 0, 0, 0)], self.cout_p, self.cout_p, hi)
            if hi:
                _hi_launch(g, vt, 1, self.cout_p, wb["shuffle"], self.kh * self.kw * self.cin, (n, oh, ow), out, wk, bias, act,
                           su=(self.kh, self.kw), cout_per_sub=self.cin)
            else:
                fprop(g, vt, 1, self.cout_p, wb["shuffle"], self.kh * self.kw * self.cin, (n, oh, ow), out, out_kind=kind,
                      su=(self.kh, self.kw), cout_per_sub=self.cin, bias=bias, act=act, wk=wk, stats=stats)
            return out
        if "s1" in wb:
            taps = [(0, self.pw - q, 0, self.ph - r) for r in range(self.kh) for q in range(self.kw)]
            vt, wk = _vtaps(taps, self.cout_p, self.cout_p, hi)
            if hi:
                _hi_launch(g, vt, 1, self.cout_p, wb["s1"], self.cin, (n, h, w), out, wk, bias, act)
            else:
                fprop(g, vt, 1, self.cout_p, wb["s1"], self.cin, (n, h, w), out, out_kind=kind, bias=bias, act=act, wk=wk,
                      stats=stats)
            return out
        if hi:
            out.zero_()
        for (a, b), (mat, taps) in wb["parity"].items():
            vt, wk = _vtaps(taps, self.cout_p, self.cout_p, hi)
            if hi:
                ksteps = len(vt) * (self.cout_p // 64)
                fprop(g, vt, 1, self.cout_p, mat, self.cin, (n, h // 2, w // 2), out, out_kind=2, su=(2, 2), sub0=(a, b),
                      cout_per_sub=self.cin, bias=bias, act=0, wk=wk, ksplit=max(1, -(-ksteps // HI_STEPS_PER_SPLIT)))
            else:
                fprop(g, vt, 1, self.cout_p, mat, self.cin, (n, h // 2, w // 2), out, out_kind=kind, su=(2, 2), sub0=(a, b),
                      cout_per_sub=self.cin, bias=bias, act=act, wk=wk)
        if hi and act:
            ops.act_fwd_(out, act)
        return out

    def backward_weight(self, dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """fp32 gradient in OIHW order as a (permuted) view [cout, cin, kh, kw] of the kernel's [cout][r][q][ci] result."""
        n, oh, ow, _ = dy.shape
        _, h, w, _ = x.shape
        hi = dy.dtype == F32
        if self.flat:
            assert x.is_contiguous()
            k = h * w * self.cin
            xf = x.view(n, 1, 1, k)
            dw = torch.empty((self.cout, k), dtype=F32, device=x.device)
            if hi:
                pairs = [(pa * self.cout_p, pb * k) for (pa, pb) in PAIRS]
                wgrad(ops.split3(dy), self.cout, ops.split3(xf), [(0, 0, 0, 0)], 1, k, (n, 1, 1), dw, pairs=pairs,
                      ksplit=_hi_wgrad_split(n, 1, 1))
            else:
                wgrad(dy, self.cout, xf, [(0, 0, 0, 0)], 1, k, (n, 1, 1), dw)
            return dw.view(self.cout, self.kh, self.kw, self.cin).permute(0, 3, 1, 2)
        xs = ops.split3(x) if hi else pad_channels(x, self.cin_p)
        gs = ops.split3(dy) if hi else dy
        taps = conv_taps(self.kh, self.kw, self.s, self.ph, self.pw, xs.stride(2))
        pairs = [(pa * self.cout_p, pb * self.cin_p) for (pa, pb) in PAIRS] if hi else None
        dw = torch.empty((self.cout, len(taps) * self.cin_p), dtype=F32, device=x.device)
        wgrad(gs, self.cout, xs, taps, self.s, self.cin_p, (n, oh, ow), dw, pairs=pairs,
              ksplit=_hi_wgrad_split(n, oh, ow) if hi else 0)
        return dw.view(self.cout, self.kh, self.kw, self.cin_p)[..., :self.cin].permute(0, 3, 1, 2)
