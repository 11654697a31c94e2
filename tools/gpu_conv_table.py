"""Per-shape table of the tensor-core launches of one eager training step (dev tool): which layers sit furthest
below the tensor roofline.  Uses the CUDA-event hook of vae_gan_mark_b200.conv (the same one bench.py's roofline
object uses), so the times are warm, on the launching stream, launch by launch.

    python tools/gpu_conv_table.py v2_128|unet_256|... [batch]
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from vae_gan_mark_b200 import conv  # noqa: E402
from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer  # noqa: E402


def main():
    wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "v2_128"])
    if len(sys.argv) > 2:
        wl["batch"] = int(sys.argv[2])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    G, D = bench.build_models(wl, dev)
    trainer = VAEGANTrainer(G, D, LossWeights.for_family(wl["family"], perceptual=False))
    B, h, w = wl["batch"], wl["h"], wl["w"]
    gen = torch.Generator(device=dev).manual_seed(1)
    batch = (torch.rand(B, 3, h, w, device=dev, generator=gen), torch.rand(B, 3, h, w, device=dev, generator=gen),
             (torch.rand(B, 1, h, w, device=dev, generator=gen) > 0.5).float())
    texts = [bench.TEXTS[i % len(bench.TEXTS)] for i in range(B)]
    for _ in range(3):
        trainer.step(*batch, texts)
    torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0.0, 0])
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ms = 0.0
    for _ in range(reps):
        conv.PROFILE = []
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.12 * 1.9e9))      # let the host run ahead: a CUDA-event pair also counts the wait for a launch to arrive
        e0.record()
        trainer.step(*batch, texts)
        e1.record()
        torch.cuda.synchronize()
        step_ms += e0.elapsed_time(e1)
        for kind, key, flops, a, b in conv.PROFILE:
            r = agg[(kind, key)]
            r[0] += a.elapsed_time(b)
            r[1] += flops
            r[2] += 1
    conv.PROFILE = None
    tot_ms = sum(v[0] for v in agg.values()) / reps
    tot_fl = sum(v[1] for v in agg.values()) / reps
    print(f"{wl['name']} batch {B}: eager step incl. a 63 ms pre-sleep {step_ms / reps:.2f} ms, tensor-core launches {tot_ms:.2f} ms, "
          f"{tot_fl / tot_ms / 1e9:.0f} TFLOP/s weighted")
    print(f"{'kind':6s} {'pixels (n,h,w)':>18s} {'N':>6s} {'K':>6s} {'#/step':>6s} {'ms/step':>8s} {'TFLOP/s':>8s} {'ms lost vs 1400':>15s}")
    rows = []
    for (kind, key), (ms, fl, n) in agg.items():
        ms, fl, n = ms / reps, fl / reps, n / reps
        rows.append((ms - fl / 1.4e12, kind, key, n, ms, fl))
    for lost, kind, key, n, ms, fl in sorted(rows, reverse=True):
        m, ng, k = key
        print(f"{kind:6s} {str(m):>18s} {ng:6d} {k:6d} {n:6.0f} {ms:8.3f} {fl / ms / 1e9:8.0f} {lost:15.3f}")


if __name__ == "__main__":
    main()
