"""Run a few eager training steps of the bench workload; the LAST one is bracketed by cudaProfilerStart/Stop so
`ncu --profile-from-start off` records exactly one step (dev tool for the launch list under profiles/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "v2_128"]
    if len(sys.argv) > 2 and sys.argv[2] == "dedup":
        from vae_gan_mark_b200 import modules as M
        M.FILM_ROW_DEDUP = True
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
    torch.cuda.cudart().cudaProfilerStart()
    trainer.step(*batch, texts)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("ok")


if __name__ == "__main__":
    main()
