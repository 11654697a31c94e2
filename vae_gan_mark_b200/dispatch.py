"""The kernel launches of this package as ``torch.library`` custom ops (namespace ``vaegan``).

north_star: "Python/PyTorch host code calling hand-written sm_100a CUDA kernels through a thin C-ABI extension registered
as torch.library custom ops".  Every function that enqueues a kernel of ``libvaegan_b200.so`` is registered with the
PyTorch dispatcher under the name of the C entry point it wraps -- ``torch.ops.vaegan.vg_conv_fprop``,
``torch.ops.vaegan.vg_norm_backward``, ``torch.ops.vaegan.vg_multi_adam`` ... -- with a schema that names the tensors a
launch writes (``Tensor(a!)``), and the package itself calls its kernels THROUGH those ops: ``ops.py`` / ``conv.py`` /
``train.FusedAdam`` functions are the dispatcher entries (``launch_op`` below returns a caller of the registered
``OpOverload``, not the Python function), the ``torch.autograd.Function``s of ``layers.py`` compose them.  The ops are
launch-level (mutable outputs, caller-provided buffers, channel-slice views): state that a functional op could not carry
-- cached bf16 weight operands, spectral-norm snapshots, concat buffers, statistics buffers -- stays with the caller and
enters as explicit tensor arguments.  ``torch_ops.py`` additionally offers functional, differentiable ops
(``vaegan::conv2d`` ...) built on the same launches.

No autograd formula is registered for the launch-level ops (they run inside ``autograd.Function.forward / backward`` or
under ``no_grad``); one implementation serves every dispatch key that reaches it (``CompositeExplicitAutograd``) and
raises for non-CUDA tensors -- there is no CPU fallback.
"""
from __future__ import annotations

import functools
import re
from typing import Callable, Dict

import torch

_FRAGMENT = torch.library.Library("vaegan", "FRAGMENT")
REGISTERED: Dict[str, str] = {}          # op name -> schema (tests/test_cabi_cpu.py checks them against the header)
_COERCE = {"int": int, "bool": bool, "float": float}


def launch_op(schema: str) -> Callable:
    """Register ``fn`` as ``torch.ops.vaegan.<name>`` (``schema`` = "name(args) -> returns") and return a function that
    calls the registered op.  Scalar arguments are coerced to the schema's type (callers pass ``int`` flags for ``bool``
    parameters and the like)."""
    name = schema[:schema.index("(")].strip()
    arg_src = schema[schema.index("(") + 1:schema.index(") ->")]
    kinds = []
    for a in [s.strip() for s in arg_src.split(",") if s.strip() and s.strip() != "*"]:
        typ, ident = a.split("=")[0].strip().rsplit(" ", 1)
        kinds.append((ident, _COERCE.get(typ)))
    by_name = dict(kinds)

    def deco(fn: Callable) -> Callable:
        _FRAGMENT.define(schema)
        _FRAGMENT.impl(name, fn, "CompositeExplicitAutograd")
        op = getattr(torch.ops.vaegan, name).default
        REGISTERED[name] = schema

        @functools.wraps(fn)
        def call(*args, **kwargs):
            if any(k is not None for _, k in kinds):
                args = tuple(k(v) if (k is not None and v is not None) else v for v, (_, k) in zip(args, kinds))
                for key, v in kwargs.items():
                    k = by_name.get(key)
                    if k is not None and v is not None:
                        kwargs[key] = k(v)
            return op(*args, **kwargs)

        call.impl = fn
        call.op = op
        return call

    return deco


def registered_entry_points():
    """C entry points that are reachable as torch.ops.vaegan.* (op name == symbol name)."""
    return sorted(n for n in REGISTERED if re.match(r"vg_", n))
