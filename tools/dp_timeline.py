#!/usr/bin/env python
"""Where do the gradient all-reduces of the data-parallel step run relative to the backward kernels?  (dev tool)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_timeline.py [workload] [nccl_sms]

Rank 0 profiles a few graph replays with the CUDA activity tracer and prints, for one step: the step span, every NCCL
kernel (start offset, duration, how much of it is covered by other kernels running at the same time) and the kernels
that were running when it started / ended."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from vae_gan_mark_b200.parallel import DataParallelReducer  # noqa: E402
from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "v2_128"]
    nccl_sms = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G, D = bench.build_models(wl, dev)
    red = DataParallelReducer(world)
    red.broadcast_parameters(list(G.parameters()) + list(D.parameters()) + list(G.buffers()) + list(D.buffers()))
    tr = VAEGANTrainer(G, D, LossWeights.for_family(wl["family"], perceptual=False), grad_hook=red.hook)
    red.install_hooks(tr.opt_G.params, tr.opt_D.params)
    if nccl_sms:
        tr.backward_sm_limit = torch.cuda.get_device_properties(dev).multi_processor_count - nccl_sms
    B, h, w = wl["batch"], wl["h"], wl["w"]
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    batch = (torch.rand(B, 3, h, w, device=dev, generator=gen), torch.rand(B, 3, h, w, device=dev, generator=gen),
             (torch.rand(B, 1, h, w, device=dev, generator=gen) > 0.5).float())
    texts = [bench.TEXTS[i % len(bench.TEXTS)] for i in range(B)]
    tr.capture(*batch, texts)
    for _ in range(3):
        tr.replay(*batch)
    dist.barrier(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            tr.replay(*batch)
        torch.cuda.synchronize()
    dist.barrier()
    if rank != 0:
        dist.destroy_process_group()
        return
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start],
                 key=lambda e: e.time_range.start)
    # second replay only
    t_first = evs[0].time_range.start
    gaps = [(evs[i + 1].time_range.start - evs[i].time_range.end, i) for i in range(len(evs) - 1)]
    cut = max(gaps)[1]            # the largest gap separates the two replays
    step = evs[cut + 1:]
    t0, t1 = step[0].time_range.start, max(e.time_range.end for e in step)
    print(f"{wl['name']} world {world} nccl_sms {nccl_sms}: step span {(t1 - t0) / 1e3:.2f} ms, {len(step)} kernels")
    others = [e for e in step if "nccl" not in e.name.lower()]
    for e in step:
        if "nccl" not in e.name.lower():
            continue
        s, f = e.time_range.start, e.time_range.end
        cov = sum(max(0, min(f, o.time_range.end) - max(s, o.time_range.start)) for o in others)
        at_start = [o.name[:40] for o in others if o.time_range.start <= s < o.time_range.end]
        print(f"  nccl  +{(s - t0) / 1e3:7.2f} ms  dur {(f - s) / 1e3:6.2f} ms  other kernels busy during it {cov / 1e3:6.2f} ms  running at start: {at_start[:2]}")
    busy = sum(o.time_range.end - o.time_range.start for o in others)
    print(f"  non-NCCL kernel time (sum) {busy / 1e3:.2f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
