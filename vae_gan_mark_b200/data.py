"""GPU side of the reference's data path: ``perspective_crop`` + ``T.ToTensor()`` (vae-gan.py:163-188, 275-281; the
same function in all five scripts) on page images that already sit in device memory.

``cv2.getPerspectiveTransform`` + ``cv2.warpPerspective(INTER_LINEAR, BORDER_REPLICATE)`` on 8-bit images is integer /
byte arithmetic; ``vg_warp.cu`` restates OpenCV's fixed-point algorithm, so the patch bytes -- and therefore the float
tensors the training step consumes -- are identical to the reference's.  The 3x3 matrix is a handful of double
operations per patch and is computed on the host by the library (``vg_perspective_crop_matrix``), operation for
operation as OpenCV does.  No CPU fallback: the image must be a CUDA uint8 tensor.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import torch

from . import _lib
from .ops import stream


def perspective_crop_matrix(bbox, out_shape: Tuple[int, int]):
    """Inverse (destination -> source) map of ``perspective_crop`` for ``bbox`` (4 points) and out_shape = (W, H), as a
    ctypes array of 9 doubles.  Raises VgError for a degenerate quadrilateral (the reference's function catches the
    OpenCV error and returns a black patch, vae-gan.py:185-188; the caller decides)."""
    pts = [float(v) for p in bbox for v in p]
    if len(pts) != 8:
        raise ValueError(f"bbox must have the shape (4, 2), got {bbox!r}")
    w, h = out_shape
    box = (C.c_float * 8)(*pts)
    minv = (C.c_double * 9)()
    _lib.call("vg_perspective_crop_matrix", box, int(w), int(h), minv)
    return minv


def _crop(image_u8: torch.Tensor, bbox, out_shape: Tuple[int, int], to_tensor: bool, host_twin: bool = False) -> torch.Tensor:
    """Argument marshalling shared by the CUDA entry point and (tests only, ``host_twin``) the host twin of its per-pixel
    code, which takes the same arguments minus the stream."""
    img = image_u8 if image_u8.dim() == 3 else image_u8.unsqueeze(2)
    if img.stride(2) != 1 or img.stride(1) != img.shape[2]:
        img = img.contiguous()
    sh, sw, ch = img.shape
    w, h = out_shape
    minv = perspective_crop_matrix(bbox, out_shape)
    if to_tensor:
        out = torch.empty((ch, h, w), dtype=torch.float32, device=img.device)
        u8, chw = C.c_void_p(0), C.c_void_p(out.data_ptr())
    else:
        out = torch.empty((h, w, ch), dtype=torch.uint8, device=img.device)
        u8, chw = C.c_void_p(out.data_ptr()), C.c_void_p(0)
    args = (C.c_void_p(img.data_ptr()), sh, sw, ch, C.c_longlong(img.stride(0)), minv, h, w, u8, chw)
    if host_twin:
        _lib.call("vg_debug_warp_perspective_host", *args)
    else:
        _lib.call("vg_warp_perspective_u8", *args, stream())
    if not to_tensor and image_u8.dim() == 2:
        out = out[:, :, 0]
    return out


def perspective_crop(image_u8: torch.Tensor, bbox, out_shape: Tuple[int, int], to_tensor: bool = True) -> torch.Tensor:
    """``T.ToTensor()(perspective_crop(image, bbox, out_shape))`` for a CUDA uint8 image (H, W, C) or (H, W):
    float32 (C, H_out, W_out) in [0, 1]; with ``to_tensor=False`` the uint8 patch (H_out, W_out[, C]) itself."""
    if not (image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.dim() in (2, 3)):
        raise RuntimeError("perspective_crop needs a CUDA uint8 image (H, W[, C]); there is no CPU fallback")
    return _crop(image_u8, bbox, out_shape, to_tensor)


def crop_batch(images: Sequence[torch.Tensor], bboxes: Sequence, out_shape: Tuple[int, int]) -> torch.Tensor:
    """One training batch of patches, (B, C, H, W) float32: what the reference's Dataset + default_collate hand to the
    step (vae-gan.py:268-283, 291-297), with the page images already resident on the device."""
    patches = [perspective_crop(img, box, out_shape) for img, box in zip(images, bboxes)]
    return torch.stack(patches, 0)
