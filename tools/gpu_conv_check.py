"""Bring-up check of the tcgen05 conv kernels against torch (cuDNN) on the GPU box.

Development tool, not a test: each case runs in its own subprocess so that a device-side trap in one
case does not poison the CUDA context of the others.  Usage:  python tools/gpu_conv_check.py [case ...]
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)), float((a - b).abs().max())


def rnd(*shape, scale=1.0):
    import torch
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def case_gemm(cin=64, cout=128, n=2, h=16, w=16, bn=0):
    import torch
    from vae_gan_mark_b200 import conv
    x = rnd(n, h, w, cin)
    wt = rnd(cout, cin, scale=cin ** -0.5)
    out = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    conv.fprop(x, [(0, 0, 0, 0)], 1, cin, wt, cout, (n, h, w), out, force_bn=bn)
    torch.cuda.synchronize()
    ref = x.float() @ wt.float().t()
    return rel_err(out.float(), ref)


def case_conv(cin=64, cout=64, n=2, h=16, w=16, k=3, s=1, p=1, bn=0, bias=False, act=0, kind=0, ksplit=0):
    import torch
    import torch.nn.functional as F
    from vae_gan_mark_b200 import conv
    x = rnd(n, h, w, cin)
    wt = rnd(cout, k, k, cin, scale=(cin * k * k) ** -0.5)   # [co][r][q][ci]
    b = torch.randn(cout, device="cuda") if bias else None
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    dt = torch.bfloat16 if kind == 0 else torch.float32
    out = torch.zeros(n, oh, ow, cout, dtype=dt, device="cuda")
    taps = conv.conv_taps(k, k, s, p, p, x.stride(2))
    conv.fprop(x, taps, s, cin, wt.view(cout, -1), cout, (n, oh, ow), out, bias=b, act=act,
               out_kind=kind, ksplit=ksplit, force_bn=bn)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), b, stride=s, padding=p)
    if act == 1:
        ref = ref.relu()
    elif act == 2:
        ref = F.leaky_relu(ref, 0.2)
    return rel_err(out.float(), ref.permute(0, 2, 3, 1))


def case_shuffle(cin=128, cout=64, n=2, h=8, w=8):
    """ConvTranspose2d(k=2, s=2) forward written into channels [cout, 2*cout) of a wider buffer."""
    import torch
    import torch.nn.functional as F
    from vae_gan_mark_b200 import conv
    x = rnd(n, h, w, cin)
    wt = rnd(cin, cout, 2, 2, scale=cin ** -0.5)             # IOHW
    b = torch.randn(cout, device="cuda")
    wf = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous()   # [(a,b,co)][ci]
    out = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16, device="cuda")
    conv.fprop(x, [(0, 0, 0, 0)], 1, cin, wf, 4 * cout, (n, h, w), out[..., cout:], su=(2, 2), cout_per_sub=cout, bias=b)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2).permute(0, 2, 3, 1)
    e = rel_err(out[..., cout:].float(), ref)
    assert float(out[..., :cout].float().abs().max()) == 0.0, "wrote outside the channel slice"
    return e


def case_wgrad(cin=64, cout=64, n=2, h=16, w=16, k=3, s=1, p=1, bn=0, ksplit=0):
    import torch
    from vae_gan_mark_b200 import conv
    x = rnd(n, h, w, cin)
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    g = rnd(n, oh, ow, cout)
    dw = torch.full((cout, k * k * cin), 7.0, dtype=torch.float32, device="cuda")
    taps = conv.conv_taps(k, k, s, p, p, x.stride(2))
    conv.wgrad(g, cout, x, taps, s, cin, (n, oh, ow), dw, ksplit=ksplit, force_bn=bn)
    torch.cuda.synchronize()
    xin = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wref = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
    y = torch.nn.functional.conv2d(xin, wref, None, stride=s, padding=p)
    y.backward(g.float().permute(0, 3, 1, 2))
    ref = wref.grad.permute(0, 2, 3, 1).reshape(cout, -1)     # [co][(r,q,ci)]
    return rel_err(dw, ref)


def case_perf(kind="fprop", cin=512, cout=512, n=8, h=128, w=128, k=3, iters=5):
    import torch
    from vae_gan_mark_b200 import conv
    x = rnd(n, h, w, cin)
    taps = conv.conv_taps(k, k, 1, k // 2, k // 2, x.stride(2))
    if kind == "fprop":
        wt = rnd(cout, k * k * cin, scale=0.02)
        out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
        fn = lambda: conv.fprop(x, taps, 1, cin, wt, cout, (n, h, w), out)
    else:
        g = rnd(n, h, w, cout)
        dw = torch.empty(cout, k * k * cin, dtype=torch.float32, device="cuda")
        fn = lambda: conv.wgrad(g, cout, x, taps, 1, cin, (n, h, w), dw)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * n * h * w * cout * cin * k * k
    return ms, flops / ms / 1e9   # ms, TFLOP/s


CASES = {
    "gemm_64x128": (case_gemm, {}),
    "gemm_bn64": (case_gemm, dict(bn=64)),
    "gemm_k256_n256": (case_gemm, dict(cin=256, cout=256)),
    "gemm_k512_n512_big": (case_gemm, dict(cin=512, cout=512, n=4, h=32, w=32)),
    "gemm_n192_edge": (case_gemm, dict(cin=128, cout=192)),
    "gemm_m_edge": (case_gemm, dict(cin=64, cout=64, n=3, h=5, w=7)),
    "conv3x3": (case_conv, {}),
    "conv3x3_c256": (case_conv, dict(cin=256, cout=256, h=8, w=8, n=4)),
    "conv3x3_bias_relu": (case_conv, dict(bias=True, act=1)),
    "conv3x3_f32out": (case_conv, dict(kind=1)),
    "conv3x3_splitk": (case_conv, dict(cin=256, kind=2, ksplit=4)),
    "conv3x3_w28": (case_conv, dict(cin=64, cout=128, n=2, h=4, w=28)),
    "conv4x4s2": (case_conv, dict(cin=64, cout=128, k=4, s=2, p=1, bias=True, act=2)),
    "conv4x4s1p1": (case_conv, dict(cin=128, cout=64, k=4, s=1, p=1, h=8, w=8)),
    "conv2x2s2": (case_conv, dict(cin=128, cout=64, k=2, s=2, p=0)),
    "conv3x3s2": (case_conv, dict(cin=128, cout=256, k=3, s=2, p=1)),
    "convT2x2_shuffle": (case_shuffle, {}),
    "wgrad3x3": (case_wgrad, {}),
    "wgrad3x3_c128": (case_wgrad, dict(cin=128, cout=128)),
    "wgrad3x3_c256_big": (case_wgrad, dict(cin=256, cout=256, n=4, h=32, w=32)),
    "wgrad1x1": (case_wgrad, dict(cin=128, cout=256, k=1, p=0)),
    "wgrad4x4s2": (case_wgrad, dict(cin=64, cout=128, k=4, s=2, p=1)),
    "wgrad_ksplit1": (case_wgrad, dict(ksplit=1)),
    "wgrad_bn64": (case_wgrad, dict(bn=64)),
    "perf_fprop_film4": (case_perf, dict(kind="fprop")),
    "perf_wgrad_film4": (case_perf, dict(kind="wgrad")),
    "perf_fprop_c64": (case_perf, dict(kind="fprop", cin=64, cout=64, n=16)),
    "perf_fprop_film4_b64": (case_perf, dict(kind="fprop", n=64, iters=3)),
    "perf_wgrad_film4_b64": (case_perf, dict(kind="wgrad", n=64, iters=3)),
    "perf_fprop_1x1_k256": (case_perf, dict(kind="fprop", cin=256, cout=512, n=16, k=1)),
    "perf_fprop_1x1_k512": (case_perf, dict(kind="fprop", cin=512, cout=256, n=16, k=1)),
    "perf_wgrad_c64": (case_perf, dict(kind="wgrad", cin=64, cout=64, n=16)),
}


def run_group(names):
    import torch
    for nme in names:
        torch.manual_seed(0)
        fn, kw = CASES[nme]
        t = time.time()
        try:
            res = fn(**kw)
            print("RESULT " + json.dumps({"case": nme, "res": res, "sec": round(time.time() - t, 2)}), flush=True)
        except Exception as e:  # host-side errors; a device trap also lands here and poisons the rest of the group
            print("RESULT " + json.dumps({"case": nme, "fail": repr(e)[:300]}), flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--group":
        run_group(sys.argv[2:])
        return
    names = sys.argv[1:] or list(CASES)
    groups = {}
    for nme in names:
        groups.setdefault(nme.split("_")[0].rstrip("0123456789x"), []).append(nme)
    for g, members in groups.items():
        try:
            p = subprocess.run([sys.executable, __file__, "--group"] + members, capture_output=True, text=True,
                               timeout=300)
            for l in p.stdout.splitlines():
                if l.startswith("RESULT "):
                    print(l[7:], flush=True)
            if p.returncode != 0:
                print(json.dumps({"group": g, "rc": p.returncode, "tail": (p.stdout + p.stderr)[-800:]}), flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"group": g, "fail": "timeout"}), flush=True)


if __name__ == "__main__":
    main()
