"""Halo-mode fprop (one activation halo per 64-channel chunk, nine shifted UMMA descriptors) against the per-tap
path: correctness for both descriptor base_offset conventions, then timing (dev tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_gan_mark_b200 import conv  # noqa: E402
from vae_gan_mark_b200.conv import ConvLinear  # noqa: E402


def run(op, x, wf, mode):
    conv.HALO_MODE = mode
    try:
        y = op.forward(x, wf)
        torch.cuda.synchronize()
        return y
    finally:
        conv.HALO_MODE = -1


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    good = None
    for (n, h, w, cin, cout) in [(2, 32, 24, 64, 64), (3, 40, 20, 128, 64), (2, 16, 8, 64, 128), (1, 50, 30, 64, 32)]:
        op = ConvLinear(cin, cout, 3, 3, 1, (1, 1))
        x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
        wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
        wf = op.prep_fwd(wt)
        ref = run(op, x, wf, -1).float()
        for mode in (1, 2):
            try:
                y = run(op, x, wf, mode).float()
                err = float((y - ref).norm() / ref.norm())
            except Exception as e:   # noqa: BLE001
                err = f"failed: {e}"
            print(f"shape n{n} {h}x{w} {cin}->{cout} halo mode {mode}: rel err vs per-tap path {err}", flush=True)
            if isinstance(err, float) and err < 1e-3:
                good = mode if good in (None, mode) else good
    if good is None:
        print("no halo mode reproduces the per-tap path")
        return
    print("using halo mode", good)
    for (n, h, w, cin, cout) in [(64, 128, 128, 64, 64), (64, 128, 128, 128, 64), (64, 128, 128, 64, 128), (64, 64, 64, 128, 128)]:
        op = ConvLinear(cin, cout, 3, 3, 1, (1, 1))
        x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
        wf = op.prep_fwd(torch.randn(cout, cin, 3, 3, device="cuda") * 0.05)
        out = conv.new_act(n, h, w, cout, "cuda")
        res = {}
        for mode in (-1, good):
            conv.HALO_MODE = mode
            res[mode] = timeit(lambda: op.forward(x, wf, out=out))
        conv.HALO_MODE = -1
        fl = 2.0 * n * h * w * cout * 9 * cin
        print(f"{n}x{h}x{w} {cin}->{cout}: per-tap {res[-1]:.4f} ms ({fl / res[-1] / 1e9:.0f} TF/s)   halo {res[good]:.4f} ms "
              f"({fl / res[good] / 1e9:.0f} TF/s)", flush=True)


if __name__ == "__main__":
    main()
