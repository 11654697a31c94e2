#!/usr/bin/env python
"""The two dominant tensor-core launches of the bench workload (FiLM4's 3x3 512->512 at 128x128, batch 64: forward and
weight gradient), three times each -- the target of the `ncu --set full` captures under profiles/ (dev tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_gan_mark_b200.conv import ConvLinear, new_act  # noqa: E402


def main():
    n, h, w = 64, 128, 128
    op = ConvLinear(512, 512, 3, 3, 1, (1, 1))
    g = torch.Generator().manual_seed(1)
    x = new_act(n, h, w, 512, "cuda"); x.copy_(torch.randn(n, h, w, 512, generator=g).to(torch.bfloat16))
    dy = new_act(n, h, w, 512, "cuda"); dy.copy_(torch.randn(n, h, w, 512, generator=g).to(torch.bfloat16))
    wf = op.prep_fwd((torch.randn(512, 512, 3, 3, generator=g) * 0.02).cuda())
    out = new_act(n, h, w, 512, "cuda")
    stats = torch.empty((1, 2, 512), device="cuda")
    for _ in range(3):
        op.forward(x, wf, None, 0, out=out, stats=stats)
    for _ in range(3):
        op.backward_weight(dy, x)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
