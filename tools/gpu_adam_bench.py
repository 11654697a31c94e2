"""clip_grad_norm_ + Adam over the real parameter lists of a bench workload, timed alone (dev tool).

    python tools/gpu_adam_bench.py [workload]          # VG_ADAM_ROTATE=0 for the fixed thread -> element map

Builds the workload's generator / discriminator the way the trainer does (channels_last conv weights, bf16 operand
shadows), fills every .grad with noise and times `FusedAdam.step(max_norm=1.0)` (table upload + multi_sumsq + final_sum +
adam_prepare + multi_adam) with CUDA events over 20 calls captured in one CUDA graph.  Bytes per parameter: 4 (norm pass
reads g) + 16 (p, g, m, v read) + 16 (p, m, v and the clipped g written) + 2 per shadowed parameter.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "v2_128"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    G, D = bench.build_models(wl, dev)
    trainer = VAEGANTrainer(G, D, LossWeights.for_family(wl["family"], perceptual=False))
    peak = bench.peaks()["hbm"]
    for name, opt in (("generator", trainer.opt_G), ("discriminator", trainer.opt_D)):
        for p in opt.params:
            p.grad = torch.randn_like(p) * 1e-3
        n = sum(p.numel() for p in opt.params)
        nsh = sum(s.numel() for s in opt.shadows if s is not None)
        small = sum(1 for p in opt.params if p.numel() < 65536)
        nbytes = n * 36 + nsh * 2
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                opt.step(max_norm=1.0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    opt.step(max_norm=1.0)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(json.dumps({"workload": wl["name"], "list": name, "tensors": len(opt.params), "tensors_below_64K": small,
                          "params_M": round(n / 1e6, 2), "MB": round(nbytes / 1e6, 1), "ms_clip_plus_adam": round(ms, 4),
                          "GB/s": round(nbytes / ms / 1e6, 1), "frac_of_copy_peak": round(nbytes / ms / 1e6 / peak, 3),
                          "VG_ADAM_ROTATE": os.environ.get("VG_ADAM_ROTATE", "1")}), flush=True)


if __name__ == "__main__":
    main()
