"""B200-native VAE-GAN training step (drop-in for Andrey1408/vae-gan-mark's nn.Module surface).

Hand-written sm_100a CUDA kernels (csrc/) behind a C ABI (include/vaegan_b200.h), called from
PyTorch host code.  No CPU fallback: the ops raise if libvaegan_b200.so is missing.
"""
__version__ = "0.1.0"


def set_precision(mode: str) -> None:
    """"bf16" (default): bf16 activations and tensor-core inputs, fp32 accumulation.  "fp32": fp32 activations, every
    tensor-core operand split into three bf16 planes (6x the tensor-core work) -- the fp32-tolerance parity mode."""
    from . import ops
    ops.set_precision(mode)
