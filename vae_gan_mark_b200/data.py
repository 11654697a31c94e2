"""GPU side of the reference's data path: ``perspective_crop`` + ``T.ToTensor()`` (vae-gan.py:163-188, 275-281; the
same function in all five scripts) on page images that already sit in device memory.

``perspective_unwarp`` (vae-gan.py:190-200, the inverse: the generated patch pasted back into the page with
BORDER_TRANSPARENT) is here as well, and ``crop_batch`` cuts a whole training batch of patches in one launch.

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
from ._lib import VgWarpJob
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


def _as_hwc(image_u8: torch.Tensor) -> torch.Tensor:
    img = image_u8 if image_u8.dim() == 3 else image_u8.unsqueeze(2)
    if img.stride(2) != 1 or img.stride(1) != img.shape[2]:
        img = img.contiguous()
    return img


def crop_batch(images: Sequence[torch.Tensor], bboxes: Sequence, out_shape: Tuple[int, int]) -> torch.Tensor:
    """One training batch of patches, (B, C, H, W) float32: what the reference's Dataset + default_collate hand to the
    step (vae-gan.py:268-283, 291-297), with the page images already resident on the device.  ONE kernel launch for the
    whole batch (``vg_warp_perspective_u8_batch``: a device table of jobs, one grid row per patch) writing straight into
    the batch tensor -- no per-patch launches, no ``torch.stack`` copy."""
    assert len(images) == len(bboxes) and len(images) >= 1
    w, h = out_shape
    imgs = []
    for im in images:
        if not (im.is_cuda and im.dtype == torch.uint8 and im.dim() in (2, 3)):
            raise RuntimeError("crop_batch needs CUDA uint8 images (H, W[, C]); there is no CPU fallback")
        imgs.append(_as_hwc(im))
    ch = imgs[0].shape[2]
    assert all(im.shape[2] == ch for im in imgs), "all images of a batch must have the same channel count"
    out = torch.empty((len(imgs), ch, h, w), dtype=torch.float32, device=imgs[0].device)
    jobs = (VgWarpJob * len(imgs))()
    for i, (im, box) in enumerate(zip(imgs, bboxes)):
        j = jobs[i]
        j.src, j.src_h, j.src_w, j.channels, j.src_row_bytes = im.data_ptr(), im.shape[0], im.shape[1], ch, im.stride(0)
        minv = perspective_crop_matrix(box, out_shape)
        for k in range(9):
            j.minv[k] = minv[k]
        j.out_h, j.out_w, j.dst_u8, j.dst_chw, j.transparent = h, w, None, out[i].data_ptr(), 0
    _launch_jobs(jobs, out.device)
    return out


def _launch_jobs(jobs, device) -> None:
    """Copy a host table of ``VgWarpJob`` to the device on the current stream and launch it (the table's device copy
    stays alive through the caching allocator until the stream has consumed it)."""
    nbytes = C.sizeof(jobs)
    host = torch.frombuffer(memoryview(jobs).cast("B"), dtype=torch.uint8)      # shares the ctypes memory
    dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dev.copy_(host, non_blocking=False)
    _lib.call("vg_warp_perspective_u8_batch", jobs, C.c_void_p(dev.data_ptr()), len(jobs), stream())


def perspective_unwarp_matrix(bbox, patch_shape: Tuple[int, int]):
    """Inverse map of ``perspective_unwarp`` (vae-gan.py:193-196): patch rectangle (W, H) -> ``bbox`` on the canvas."""
    pts = [float(v) for p in bbox for v in p]
    if len(pts) != 8:
        raise ValueError(f"bbox must have the shape (4, 2), got {bbox!r}")
    w, h = patch_shape
    minv = (C.c_double * 9)()
    _lib.call("vg_perspective_unwarp_matrix", (C.c_float * 8)(*pts), int(w), int(h), minv)
    return minv


def perspective_unwarp(patch_u8: torch.Tensor, bbox, canvas_shape, canvas: torch.Tensor = None, host_twin: bool = False) -> torch.Tensor:
    """``perspective_unwarp(patch, bbox, canvas_shape)`` of the reference (vae-gan.py:190-200) for a CUDA uint8 patch
    (h, w[, C]): a zero uint8 canvas of ``canvas_shape`` = (H, W[, C]) with the patch pasted through the inverse
    perspective map (INTER_LINEAR, BORDER_TRANSPARENT).  ``canvas``: paste into this existing canvas instead (e.g. the
    page image, to compose the translated patch into it)."""
    if not (patch_u8.dtype == torch.uint8 and patch_u8.dim() in (2, 3) and (patch_u8.is_cuda or host_twin)):
        raise RuntimeError("perspective_unwarp needs a CUDA uint8 patch (h, w[, C]); there is no CPU fallback")
    img = _as_hwc(patch_u8)
    ph, pw, ch = img.shape
    shape = tuple(canvas_shape)
    assert len(shape) in (2, 3) and (len(shape) == 2) == (patch_u8.dim() == 2) and (len(shape) == 2 or shape[2] == ch)
    if canvas is None:
        canvas = torch.zeros(shape, dtype=torch.uint8, device=img.device)
    assert tuple(canvas.shape) == shape and canvas.dtype == torch.uint8 and canvas.is_contiguous() and canvas.device == img.device
    minv = perspective_unwarp_matrix(bbox, (pw, ph))
    args = (C.c_void_p(img.data_ptr()), ph, pw, ch, C.c_longlong(img.stride(0)), minv, shape[0], shape[1],
            C.c_void_p(canvas.data_ptr()))
    if host_twin:
        _lib.call("vg_debug_warp_perspective_transparent_host", *args)
    else:
        _lib.call("vg_warp_perspective_u8_transparent", *args, stream())
    return canvas
