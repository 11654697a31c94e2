"""How much of a graph-replayed step is idle time between kernels?  (dev tool)
Profiles a few replays with the CUDA activity tracer and prints wall span, summed kernel time and the gap histogram."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "v2_128"]
    if len(sys.argv) > 1 and sys.argv[1] == "dedup":
        from vae_gan_mark_b200 import modules as M
        M.FILM_ROW_DEDUP = True
    dev = torch.device("cuda", 0)
    G, D = bench.build_models(wl, dev)
    tr = VAEGANTrainer(G, D, LossWeights.for_family(wl["family"], perceptual=False))
    B, h, w = wl["batch"], wl["h"], wl["w"]
    gen = torch.Generator(device=dev).manual_seed(1)
    batch = (torch.rand(B, 3, h, w, device=dev, generator=gen), torch.rand(B, 3, h, w, device=dev, generator=gen),
             (torch.rand(B, 1, h, w, device=dev, generator=gen) > 0.5).float())
    texts = [bench.TEXTS[i % len(bench.TEXTS)] for i in range(B)]
    tr.capture(*batch, texts)
    for _ in range(3):
        tr.replay(*batch)
    torch.cuda.synchronize()
    nrep = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(nrep):
            tr.replay(*batch)
        torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start],
                 key=lambda e: e.time_range.start)
    span = (evs[-1].time_range.end - evs[0].time_range.start) / 1e3
    busy = sum(e.time_range.end - e.time_range.start for e in evs) / 1e3
    gaps = []
    end = evs[0].time_range.end
    for e in evs[1:]:
        gaps.append(max(0.0, e.time_range.start - end))
        end = max(end, e.time_range.end)
    big = sorted(((g, i) for i, g in enumerate(gaps)), reverse=True)[:6]
    for g, i in big:
        print(f"gap {g:9.1f} us after #{i} {evs[i].name[:50]!r} before {evs[i + 1].name[:50]!r}")
    gaps.sort()
    print(f"replays {nrep}: kernels {len(evs)} span {span:.2f} ms busy(sum) {busy:.2f} ms gaps(sum) {sum(gaps) / 1e3:.2f} ms")
    n = len(gaps)
    print("gap us: median %.2f  p90 %.2f  p99 %.2f  max %.2f" % (gaps[n // 2], gaps[int(n * .9)], gaps[int(n * .99)], gaps[-1]))
    by = {}
    for e in evs:
        k = e.name[:60]
        d = by.setdefault(k, [0.0, 0])
        d[0] += (e.time_range.end - e.time_range.start) / 1e3
        d[1] += 1
    for k, (ms, c) in sorted(by.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"{ms:8.3f} ms {c:4d}  {k}")


if __name__ == "__main__":
    main()
