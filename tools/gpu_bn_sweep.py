"""Sweep the forced N tile for memory-bound 1x1 convs (development tool)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gan_mark_b200 import conv

def run(cin, cout, n, h, w, k, bn):
    x = (torch.randn(n, h, w, cin, device="cuda")).to(torch.bfloat16)
    taps = conv.conv_taps(k, k, 1, k // 2, k // 2, x.stride(2))
    wt = (torch.randn(cout, k * k * cin, device="cuda") * 0.02).to(torch.bfloat16)
    out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    fn = lambda: conv.fprop(x, taps, 1, cin, wt, cout, (n, h, w), out, force_bn=bn)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (x.numel() + out.numel()) * 2 / 1e9
    print(json.dumps({"cin": cin, "cout": cout, "k": k, "m": [n, h, w], "bn": bn, "ms": round(ms, 4),
                      "TF/s": round(2.0 * n * h * w * cout * cin * k * k / ms / 1e9, 1), "GB/s": round(gb / ms * 1e3, 1)}), flush=True)

for (cin, cout, k) in [(512, 256, 1), (256, 512, 1), (64, 64, 3), (128, 64, 3), (64, 64, 1)]:
    for bn in (64, 128, 256):
        if bn > max(cout, 64): continue
        run(cin, cout, 64, 128, 128, k, bn)
