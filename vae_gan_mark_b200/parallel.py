"""Data parallelism for the training step: one process per GPU, gradients averaged with bucketed NCCL
all-reduces over NVLink/NVSwitch (the reference is single-process).

The batch dimension shards naturally; the only exchange per step is
  * all-reduce(D grads)  after ``loss_D.backward()``  (2.8 M params, 11 MB fp32), and
  * all-reduce(G grads)  after ``loss_G.backward()``  (66 M params at cfg 2/4),
both pre-scaled by 1/world so the clip / Adam kernels see the global-batch mean.  BatchNorm statistics stay
per-replica (DistributedDataParallel semantics).  Buckets are filled in reverse parameter order (the order the
backward produces gradients) and launched asynchronously on NCCL's stream as they fill, overlapping with the
rest of the backward when ``install_hooks`` is used; ``hook`` waits for all of them before the optimiser runs.

Copies.  A bucket is packed with ONE fused concatenation of memory-order views.  There is no unpack: after the
all-reduce the averaged gradients stay where NCCL left them and every ``p.grad`` becomes a strided view into the
bucket (the clip / Adam kernels walk raw pointers), so the exchange costs one read + one write of the gradients
instead of two.  ``grad_dtype=torch.bfloat16`` halves the bytes on the wire (the pack casts to bf16, NCCL averages in
bf16, the unpack casts back into the fp32 gradients): meant for the 568 M-parameter configuration (BASELINE configs[4],
2.27 GB of fp32 gradients per step), where the exchange rather than the convolutions bounds small batches; replicas stay
bit-identical either way because every rank receives the same reduced values.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


class DataParallelReducer:
    def __init__(self, world_size: int, bucket_bytes: int = 32 << 20, group=None, grad_dtype: torch.dtype = torch.float32):
        assert grad_dtype in (torch.float32, torch.bfloat16)
        self.world, self.bucket_bytes, self.group, self.grad_dtype = world_size, bucket_bytes, group, grad_dtype
        self._pending: List = []
        self._states: List[Dict] = []
        self._state_of: Dict[int, Dict] = {}
        self._handles: List = []
        self.overlap = False

    # ------------------------------------------------------------------ setup
    def broadcast_parameters(self, tensors: Sequence[torch.Tensor], src: int = 0) -> None:
        """Make every replica start from rank ``src``'s parameters and buffers."""
        if self.world <= 1:
            return
        for t in tensors:
            dist.broadcast(t.data if t.is_floating_point() or t.dtype == torch.int64 else t, src, group=self.group)
        if any(t.is_cuda for t in tensors):
            from . import layers as L      # (p.data writes do not bump version counters: refresh the bf16 operand shadows)
            L.sync_all_shadows()

    def make_buckets(self, params: Sequence[torch.nn.Parameter]) -> List[List[torch.nn.Parameter]]:
        """Reverse parameter order, ~bucket_bytes each."""
        buckets, cur, size = [], [], 0
        for p in reversed(list(params)):
            cur.append(p)
            size += p.numel() * 4
            if size >= self.bucket_bytes:
                buckets.append(cur)
                cur, size = [], 0
        if cur:
            buckets.append(cur)
        return buckets

    # ------------------------------------------------------------------ reduction
    @staticmethod
    def _memory_view(g: torch.Tensor) -> torch.Tensor:
        """1-D view of a dense gradient in its own memory order (channels_last conv-weight gradients included), so that
        packing a bucket is one fused concatenation and unpacking one fused multi-tensor copy -- no per-tensor
        launches (there are ~190 gradient tensors; per-tensor copies cost more than the all-reduce itself)."""
        if g.dim() == 4 and not g.is_contiguous() and g.is_contiguous(memory_format=torch.channels_last):
            return g.permute(0, 2, 3, 1).reshape(-1)
        assert g.is_contiguous(), "gradient is neither contiguous nor channels_last"   # reshape would copy
        return g.reshape(-1)

    def _launch(self, bucket: List[torch.nn.Parameter]) -> None:
        params = [p for p in bucket if p.grad is not None]
        views = [self._memory_view(p.grad) for p in params]
        if not views:
            return
        if self.grad_dtype == torch.float32:
            flat = torch.cat(views)
        else:                                   # cast while packing: one fused multi-tensor copy
            flat = torch.empty(sum(v.numel() for v in views), dtype=self.grad_dtype, device=views[0].device)
            torch._foreach_copy_(list(flat.split([v.numel() for v in views])), views)
        if dist.get_backend(self.group) == "nccl":
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:                                   # gloo (CPU tests) has no AVG
            flat.mul_(1.0 / self.world)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((work, flat, views, params, [(p.grad.shape, p.grad.stride()) for p in params]))

    def wait(self) -> None:
        for work, flat, views, params, layouts in self._pending:
            work.wait()
            chunks = flat.split([v.numel() for v in views])
            if self.grad_dtype == torch.float32:
                # no unpack: the averaged gradient of each parameter is the bucket's slice, viewed with the shape and
                # memory order the gradient had when it was packed (OIHW or channels_last)
                for p, c, (shape, stride) in zip(params, chunks, layouts):
                    p.grad = c.as_strided(shape, stride)
            else:
                torch._foreach_copy_(views, list(chunks))
        self._pending = []

    # ------------------------------------------------------------------ overlap with the backward pass
    def install_hooks(self, *param_lists: Sequence[torch.nn.Parameter]) -> None:
        """Launch every bucket's all-reduce from autograd as soon as its last gradient has been accumulated
        (``register_post_accumulate_grad_hook``), so the exchange overlaps with the rest of the backward pass; ``hook``
        then only launches what is left (buckets with a parameter that got no gradient) and waits.  Buckets follow
        the reverse parameter order, which is the order the backward produces gradients in.  Works inside CUDA-graph
        capture: the collectives are recorded at the point of the backward where they were issued."""
        if self.world <= 1:
            return
        for params in param_lists:
            for bucket in self.make_buckets(list(params)):
                state = {"params": bucket, "ready": 0, "launched": False}
                self._states.append(state)
                for p in bucket:
                    self._state_of[id(p)] = state
                    self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.overlap = True

    def remove_hooks(self) -> None:
        """Detach the autograd hooks again (the parameters outlive the reducer when trainers are rebuilt)."""
        for h in self._handles:
            h.remove()
        self._handles, self._states, self._state_of, self.overlap = [], [], {}, False

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        st = self._state_of.get(id(p))
        if st is None or not self.overlap:
            return
        st["ready"] += 1
        if st["ready"] == len(st["params"]) and not st["launched"]:
            st["launched"] = True
            self._launch(st["params"])

    def hook(self, which: str, params: Sequence[torch.nn.Parameter]) -> None:
        """``grad_hook`` of VAEGANTrainer: called after each backward; returns with averaged gradients in place."""
        if self.world <= 1:
            return
        if self.overlap:
            mine = {id(p) for p in params}
            for st in self._states:
                if id(st["params"][0]) not in mine:
                    continue
                if not st["launched"] and st["ready"] > 0:      # a bucket some of whose parameters got no gradient
                    self._launch(st["params"])
                st["ready"], st["launched"] = 0, False
        else:
            for bucket in self.make_buckets(params):
                self._launch(bucket)
        self.wait()
