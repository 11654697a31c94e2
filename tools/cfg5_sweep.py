#!/usr/bin/env python
"""BASELINE configs[4]: vae-gan-unet.py (row-U repair) at 256x256, latent dim 512 (568 M generator parameters, 2.27 GB of
fp32 gradients per step), batch sweep on N GPUs of one node, with a roofline / communication report.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/cfg5_sweep.py --batches 8,16,32,64,128 --out gpurun_out/cfg5_8gpu.jsonl
    python tools/cfg5_sweep.py --batches 8,16,32,64,128 --out gpurun_out/cfg5_1gpu.jsonl          # single GPU

For every per-GPU batch the same captured step is timed in up to four data-parallel variants inside ONE process group:
  nocomm     the N replicas run without any exchange (compute only, under the power / thermal conditions of N busy GPUs)
  fp32       fp32 gradient buckets, all-reduces issued from autograd hooks during the backward
  bf16       bf16 gradient buckets (half the bytes on the wire)
  bf16+sms   bf16 buckets, and K SMs kept out of the persistent conv grids during the backward so NCCL runs beside them
exposed communication = variant - nocomm.  CUDA-event timing, max over ranks, one JSON line per (batch, variant).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="8,16,32,64,128")
    ap.add_argument("--variants", default="nocomm,fp32,bf16,bf16+sms")
    ap.add_argument("--workload", default="unet_256_z512")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nccl-sms", type=int, default=16)
    ap.add_argument("--bucket-mb", type=int, default=64)
    ap.add_argument("--plan", default="", help='per-batch variant lists, e.g. "8:nocomm,fp32,bf16,bf16+sms;16:nocomm,bf16+sms" '
                                               "(overrides --batches / --variants)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import bench
    from vae_gan_mark_b200.parallel import DataParallelReducer
    from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = dict(bench.WORKLOADS[args.workload])
    h, w = wl["h"], wl["w"]
    G, D = bench.build_models(wl, dev)
    n_params = sum(p.numel() for p in G.parameters())
    state0 = ({k: v.clone() for k, v in G.state_dict().items()}, {k: v.clone() for k, v in D.state_dict().items()})
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    gflop = bench.STEP_GFLOP_PER_IMG[args.workload]
    pk = bench.peaks()
    lines = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.plan:
        plan = [(int(item.split(":")[0]), item.split(":")[1].split(",")) for item in args.plan.split(";") if item]
    else:
        plan = [(int(b), args.variants.split(",")) for b in args.batches.split(",")]
    for B, variants in plan:
        variants = [v for v in variants if world > 1 or v == "nocomm"]
        gen = torch.Generator(device=dev).manual_seed(4321 + rank)
        data = [(torch.rand(B, 3, h, w, device=dev, generator=gen), torch.rand(B, 3, h, w, device=dev, generator=gen),
                 (torch.rand(B, 1, h, w, device=dev, generator=gen) > 0.5).float()) for _ in range(2)]
        texts = [bench.TEXTS[i % len(bench.TEXTS)] for i in range(B)]
        base_ms = None
        for var in variants:
            G.load_state_dict(state0[0]); D.load_state_dict(state0[1])
            reducer = None
            if var != "nocomm":
                reducer = DataParallelReducer(world, bucket_bytes=args.bucket_mb << 20,
                                              grad_dtype=torch.bfloat16 if var.startswith("bf16") else torch.float32)
            tr = VAEGANTrainer(G, D, LossWeights.for_family(wl["family"], perceptual=False),
                               grad_hook=reducer.hook if reducer else None)
            if reducer is not None:
                reducer.install_hooks(tr.opt_G.params, tr.opt_D.params)
                if var.endswith("+sms"):
                    tr.backward_sm_limit = sms - args.nccl_sms
            tr.capture(data[0][0], data[0][1], data[0][2], texts)
            for i in range(args.warmup):
                tr.replay(*data[i % 2])
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                tr.replay(*data[i % 2])
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            if var == "nocomm":
                base_ms = ms
            in_sync = None
            if reducer is not None:
                chk = torch.stack([p.detach().double().sum() for p in list(G.parameters()) + list(D.parameters())]).sum()
                lo, hi = chk.clone(), chk.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                in_sync = bool(abs(float(hi) - float(lo)) <= 1e-9 * max(1.0, abs(float(hi))))
            line = {"workload": wl["name"], "n_gpus": world, "per_gpu_batch": B, "global_batch": B * world, "variant": var,
                    "ms_per_step": ms, "images_per_s": world * B / (ms / 1e3),
                    "exposed_comm_ms": (ms - base_ms) if base_ms is not None and var != "nocomm" else None,
                    "step_algorithmic_tflops_per_gpu": gflop * B / 1e3 / (ms / 1e3),
                    "step_frac_of_sustained_peak": gflop * B / 1e3 / (ms / 1e3) / pk["tf"],
                    "generator_params": n_params, "grad_bytes_fp32": 4 * n_params, "dp_params_in_sync": in_sync,
                    "bucket_mb": args.bucket_mb, "nccl_sms": args.nccl_sms if var.endswith("+sms") else 0,
                    "gpu_launches_per_step": tr.launches_per_step}
            lines.append(line)
            if rank == 0:
                print(json.dumps(line), flush=True)
                if args.out:
                    with open(args.out, "a") as f:
                        f.write(json.dumps(line) + "\n")
            for p in list(G.parameters()) + list(D.parameters()):
                p.grad = None
            if reducer is not None:
                reducer.remove_hooks()
            del tr, reducer
            gc.collect()
            torch.cuda.empty_cache()
        del data
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
