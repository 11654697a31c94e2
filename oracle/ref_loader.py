"""Import the reference's training scripts by path (build container only).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  ``/root/reference`` does not exist on the
GPU box; nothing on the ``-m gpu`` / smoke / bench paths calls this module.

The scripts have hyphenated names, guard ``main()`` behind ``__name__ == "__main__"``
(vae-gan.py:603) and import four packages that are absent offline; those are stubbed.  Each
script also writes a hard-coded credential into ``os.environ`` at import (line 24 of every
vae-gan*.py) -- it is scrubbed right after the import and never read.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

import torch

from .models import SBERT_DIM, hash_sentence_embedding

REFERENCE_ROOT = os.environ.get("VAEGAN_REFERENCE_ROOT", "/root/reference")
SCRIPTS = {"base": "vae-gan.py", "v2": "vae-gan-v2.py", "unet": "vae-gan-unet.py",
           "oldv": "vae-gan-oldv.py", "lrsh": "vae-gan-lr-sh.py"}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, SCRIPTS["base"]))


class _StubSentenceTransformer:
    """Stands in for sentence_transformers.SentenceTransformer (vae-gan.py:93,110)."""

    def __init__(self, *a, **k):
        pass

    def get_sentence_embedding_dimension(self):
        return SBERT_DIM

    def to(self, *_a, **_k):
        return self

    def encode(self, texts, convert_to_tensor=True, device=None):
        return hash_sentence_embedding(texts)


def _install_stubs():
    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
        return sys.modules[name]

    stub("sentence_transformers", SentenceTransformer=_StubSentenceTransformer)
    stub("torchinfo", summary=lambda *a, **k: None)
    plt = stub("matplotlib.pyplot")
    stub("matplotlib", pyplot=plt)
    stub("kagglehub")


_CACHE = {}


def load(family: str, patch_wh):
    """Return the reference module for ``family`` with ``PATCH_SHAPE`` = (W, H) set."""
    if family not in _CACHE:
        _install_stubs()
        path = os.path.join(REFERENCE_ROOT, SCRIPTS[family])
        spec = importlib.util.spec_from_file_location("vaegan_reference_" + family, path)
        mod = importlib.util.module_from_spec(spec)
        saved = os.environ.get("WANDB_API_KEY")
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
        os.environ.pop("WANDB_API_KEY", None)
        if saved is not None:
            os.environ["WANDB_API_KEY"] = saved
        mod.DEVICE = "cpu"
        _CACHE[family] = mod
    mod = _CACHE[family]
    mod.PATCH_SHAPE = tuple(patch_wh)
    return mod


def build(family: str, h: int, w: int, z_ch: int = 128):
    """Construct the reference's (G, D) for a family at patch size (h, w), quietly."""
    mod = load(family, (w, h))
    with contextlib.redirect_stdout(io.StringIO()):
        if family in ("base", "lrsh"):
            G = mod.VAEGAN(in_ch=4, z_ch=z_ch, text_ch=64, out_ch=3)
        elif family == "v2":
            G = mod.VAEGAN_UNet_SpatialFiLM(in_ch_style=4, z_ch_style=z_ch)
        elif family == "unet":
            G = mod.VAEGAN_UNet_CharEmb(in_ch_for_style_encoder=4, z_ch_for_style=z_ch)
        elif family == "oldv":
            G = mod.VAEGAN_UNet_SpatialFiLM(in_ch_style=4, z_ch_style=z_ch)
        else:
            raise ValueError(family)
        D = mod.Discriminator(3)
    return mod, G, D
