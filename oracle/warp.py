"""CPU restatement of the reference's patch extraction (test infrastructure only -- never imported by the product).

``perspective_crop`` (vae-gan.py:163-188; same function in every script) cuts a quadrilateral ``bbox`` out of a page
image with ``cv2.getPerspectiveTransform`` + ``cv2.warpPerspective(INTER_LINEAR, BORDER_REPLICATE)`` and the dataset
turns the uint8 patch into a float tensor with ``T.ToTensor()`` (vae-gan.py:275-281).  The arithmetic lives in a
third-party dependency, ``opencv-python`` (requirements.txt:4, unpinned; 4.13.0 in this image).  This file restates
OpenCV's published algorithm for 8-bit images (modules/imgproc/src/imgwarp.cpp: getPerspectiveTransform,
WarpPerspectiveInvoker, remapBilinear with the fixed-point INTER_BITS = 5 / INTER_REMAP_COEF_BITS = 15 tables) in numpy
and is PINNED against cv2 itself, bit for bit, in tests/test_warp_oracle.py (random and degenerate quadrilaterals,
1- and 3-channel images, quadrilaterals that leave the image).

Integer / byte work: the bar is bit-exact.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS          # 32 sub-pixel positions per axis
BLOCK_W = 64                              # WarpPerspectiveInvoker walks the destination in blocks this wide (BLOCK_SZ^2 / 16)


def _lu_solve(a, b):
    """cv::solve(DECOMP_LU) = hal::LU64f: Gaussian elimination with partial pivoting (first largest |pivot|), the row
    updates as separate multiply and add, then back-substitution.  Plain Python floats so that every rounding is the
    reference's."""
    a = [[float(v) for v in row] for row in a]
    b = [float(v) for v in b]
    n = len(b)
    for i in range(n):
        k = i
        for j in range(i + 1, n):
            if abs(a[j][i]) > abs(a[k][i]):
                k = j
        if k != i:
            a[i], a[k] = a[k], a[i]
            b[i], b[k] = b[k], b[i]
        d = -1.0 / a[i][i]
        for j in range(i + 1, n):
            alpha = a[j][i] * d
            for c in range(i + 1, n):
                a[j][c] = a[j][c] + alpha * a[i][c]
            b[j] = b[j] + alpha * b[i]
    for i in range(n - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, n):
            s = s - a[i][c] * b[c]
        b[i] = s / a[i][i]
    return b


def get_perspective_transform(src_quad, dst_quad) -> np.ndarray:
    """cv2.getPerspectiveTransform: the 3x3 map src -> dst (M[2,2] = 1) from four point pairs.  The points are float32
    (Point2f) and the products -x*x', -y*x' ... of the system matrix are formed in float32 before they are widened."""
    s = np.asarray(src_quad, dtype=np.float32).reshape(4, 2)
    d = np.asarray(dst_quad, dtype=np.float32).reshape(4, 2)
    a = np.zeros((8, 8), dtype=np.float64)
    b = np.zeros(8, dtype=np.float64)
    for i in range(4):
        a[i, 0] = a[i + 4, 3] = s[i, 0]
        a[i, 1] = a[i + 4, 4] = s[i, 1]
        a[i, 2] = a[i + 4, 5] = 1.0
        a[i, 6], a[i, 7] = np.float32(-s[i, 0] * d[i, 0]), np.float32(-s[i, 1] * d[i, 0])
        a[i + 4, 6], a[i + 4, 7] = np.float32(-s[i, 0] * d[i, 1]), np.float32(-s[i, 1] * d[i, 1])
        b[i], b[i + 4] = d[i, 0], d[i, 1]
    return np.array(_lu_solve(a, b) + [1.0], dtype=np.float64).reshape(3, 3)


def crop_matrix(bbox, out_shape) -> np.ndarray:
    """The matrix ``perspective_crop`` hands to warpPerspective (vae-gan.py:176-178); out_shape = (W, H)."""
    w, h = out_shape
    dst = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.float32)
    return get_perspective_transform(bbox, dst)


def inverse_map(m: np.ndarray) -> np.ndarray:
    """warpPerspective inverts the matrix (no WARP_INVERSE_MAP in the reference) and walks the destination.  cv::invert
    of a 3x3 double matrix: cofactors times the reciprocal determinant, in this operation order."""
    m = np.asarray(m, dtype=np.float64).reshape(3, 3)
    S = lambda i, j: float(m[i, j])   # noqa: E731
    det = (S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0))
           + S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0)))
    d = 1.0 / det
    t = [(S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * d, (S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * d,
         (S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * d, (S(1, 2) * S(2, 0) - S(1, 0) * S(2, 2)) * d,
         (S(0, 0) * S(2, 2) - S(0, 2) * S(2, 0)) * d, (S(0, 2) * S(1, 0) - S(0, 0) * S(1, 2)) * d,
         (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0)) * d, (S(0, 1) * S(2, 0) - S(0, 0) * S(2, 1)) * d,
         (S(0, 0) * S(1, 1) - S(0, 1) * S(1, 0)) * d]
    return np.array(t, dtype=np.float64).reshape(3, 3)


def warp_coords(minv: np.ndarray, out_w: int, out_h: int):
    """Fixed-point source coordinates of every destination pixel, as WarpPerspectiveInvoker computes them: per row and
    64-wide block X0 = M0*bx + M1*y + M2, then X = cvRound((X0 + M0*x1) * (32 / (W0 + M6*x1))), clamped to int."""
    m = np.asarray(minv, dtype=np.float64).reshape(9)
    xs = np.arange(out_w, dtype=np.int64)
    bx = (xs // BLOCK_W * BLOCK_W).astype(np.float64)
    x1 = (xs % BLOCK_W).astype(np.float64)
    ys = np.arange(out_h, dtype=np.float64)[:, None]
    x0 = (m[0] * bx[None, :] + m[1] * ys) + m[2]
    y0 = (m[3] * bx[None, :] + m[4] * ys) + m[5]
    w0 = (m[6] * bx[None, :] + m[7] * ys) + m[8]
    w = w0 + m[6] * x1[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(w != 0.0, INTER_TAB_SIZE / w, 0.0)
    lo, hi = float(np.iinfo(np.int32).min), float(np.iinfo(np.int32).max)
    fx = np.maximum(lo, np.minimum(hi, (x0 + m[0] * x1[None, :]) * w))
    fy = np.maximum(lo, np.minimum(hi, (y0 + m[3] * x1[None, :]) * w))
    return np.rint(fx).astype(np.int64), np.rint(fy).astype(np.int64)      # cvRound: half to even


def warp_perspective_u8(img: np.ndarray, m: np.ndarray, out_shape, transparent_into: np.ndarray = None) -> np.ndarray:
    """cv2.warpPerspective(img, m, (W, H), flags=INTER_LINEAR, borderMode=BORDER_REPLICATE) for uint8 images.
    ``transparent_into`` = a (H, W[, C]) uint8 canvas: borderMode=BORDER_TRANSPARENT with ``dst=canvas`` instead.
    Pinned against cv2 4.13 (tests/test_warp_oracle.py): a destination pixel is written iff the integer part of its source
    coordinate lies inside the source (0 <= sx <= sw - 1 and 0 <= sy <= sh - 1), with the same value BORDER_REPLICATE
    gives (the +1 neighbours of the last row / column are clamped); every other canvas pixel is left untouched."""
    assert img.dtype == np.uint8
    out_w, out_h = out_shape
    src = img if img.ndim == 3 else img[:, :, None]
    sh, sw = src.shape[:2]
    X, Y = warp_coords(inverse_map(m), out_w, out_h)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)          # stored as short
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    ax, ay = X & (INTER_TAB_SIZE - 1), Y & (INTER_TAB_SIZE - 1)
    x0, x1 = np.clip(sx, 0, sw - 1), np.clip(sx + 1, 0, sw - 1)      # BORDER_REPLICATE
    y0, y1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    # bilinear table of 15-bit weights: (32-ay)(32-ax) * 32 etc. are exact, so they already sum to 1 << 15
    w00 = ((INTER_TAB_SIZE - ay) * (INTER_TAB_SIZE - ax))[..., None]
    w01 = ((INTER_TAB_SIZE - ay) * ax)[..., None]
    w10 = (ay * (INTER_TAB_SIZE - ax))[..., None]
    w11 = (ay * ax)[..., None]
    p = src.astype(np.int64)
    acc = p[y0, x0] * w00 + p[y0, x1] * w01 + p[y1, x0] * w10 + p[y1, x1] * w11
    out = ((acc * 32 + (1 << 14)) >> 15).astype(np.uint8)
    if transparent_into is not None:
        canvas = transparent_into if transparent_into.ndim == 3 else transparent_into[:, :, None]
        assert canvas.shape == out.shape and canvas.dtype == np.uint8
        inside = (sx >= 0) & (sx <= sw - 1) & (sy >= 0) & (sy <= sh - 1)
        out = np.where(inside[..., None], out, canvas)
    return out if img.ndim == 3 else out[:, :, 0]


def perspective_crop(img: np.ndarray, bbox, out_shape) -> np.ndarray:
    """vae-gan.py:163-188 on a uint8 array (H, W[, C]); out_shape = (W, H).  Returns the uint8 patch."""
    return warp_perspective_u8(img, crop_matrix(bbox, out_shape), out_shape)


def unwarp_matrix(bbox, patch_shape) -> np.ndarray:
    """The matrix ``perspective_unwarp`` hands to warpPerspective (vae-gan.py:193-196): patch rectangle -> bbox;
    patch_shape = (W, H) of the patch."""
    w, h = patch_shape
    src = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.float32)
    return get_perspective_transform(src, np.asarray(bbox, dtype=np.float32).reshape(4, 2))


def perspective_unwarp(patch: np.ndarray, bbox, canvas_shape) -> np.ndarray:
    """vae-gan.py:190-200: paste a (generated) uint8 patch (h, w[, C]) back into a zero canvas of ``canvas_shape`` =
    (H, W[, C]) through the inverse perspective map, borderMode=BORDER_TRANSPARENT."""
    canvas = np.zeros(canvas_shape, dtype=np.uint8)
    h, w = patch.shape[:2]
    return warp_perspective_u8(patch, unwarp_matrix(bbox, (w, h)), (canvas_shape[1], canvas_shape[0]), transparent_into=canvas)


def to_tensor(patch: np.ndarray) -> np.ndarray:
    """T.ToTensor() (vae-gan.py:280-281): uint8 HWC -> float32 CHW in [0, 1] (a float32 division by 255)."""
    p = patch if patch.ndim == 3 else patch[:, :, None]
    return (p.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def fixture_inputs():
    """Deterministic page / mask images and quadrilaterals of tests/golden/warp_crop.npz (outputs of the reference's own
    ``perspective_crop`` + ``T.ToTensor()``, recorded by tests/golden/make_golden.py); inputs are re-derived, not stored."""
    rng = np.random.default_rng(2024)
    page = rng.integers(0, 256, size=(96, 160, 3), dtype=np.uint8)
    mask = rng.integers(0, 256, size=(96, 160), dtype=np.uint8)
    boxes = [[[10, 12], [140, 8], [150, 70], [6, 80]], [[30.5, 20.25], [120.75, 30.5], [110.25, 60.5], [25.5, 55.75]],
             [[-20, -10], [100, 5], [90, 50], [-15, 60]], [[60, 40], [200, 30], [210, 120], [55, 110]],
             [[5, 5], [68, 5], [68, 36], [5, 36]], [[150, 10], [20, 15], [25, 85], [155, 90]]]
    return page, mask, boxes


def unwarp_fixture_patches():
    """Deterministic patches of tests/golden/warp_unwarp.npz (outputs of the reference's own ``perspective_unwarp`` on the
    quadrilaterals of ``fixture_inputs``, recorded by tests/golden/make_golden.py warp_unwarp)."""
    rng = np.random.default_rng(2025)
    return {"rgb_64x448": rng.integers(0, 256, size=(64, 448, 3), dtype=np.uint8),
            "gray_32x64": rng.integers(0, 256, size=(32, 64), dtype=np.uint8)}


def unwarp_cases(rng, n):
    """(patch, bbox, canvas_shape): 1-4 channel patches down to 2 x 2, quadrilaterals partly off the canvas, rotated corner
    order (mirrored pastes)."""
    for trial in range(n):
        ch = [None, 3, 1, 3, 4, 2][trial % 6]
        h, w = int(rng.integers(2, 70)), int(rng.integers(2, 200))
        patch = rng.integers(0, 256, size=(h, w) if ch is None else (h, w, ch), dtype=np.uint8)
        H, W = int(rng.integers(60, 200)), int(rng.integers(80, 300))
        cx, cy = rng.uniform(20, W - 20), rng.uniform(20, H - 20)
        sx, sy = rng.uniform(10, W / 2), rng.uniform(8, H / 2)
        bbox = (np.array([[cx - sx, cy - sy], [cx + sx, cy - sy], [cx + sx, cy + sy], [cx - sx, cy + sy]], dtype=np.float32)
                + rng.uniform(-6, 6, size=(4, 2)).astype(np.float32))
        if trial % 5 == 0:
            bbox -= 30
        if trial % 7 == 0:
            bbox = bbox[[1, 2, 3, 0]]
        yield patch, bbox, ((H, W) if ch is None else (H, W, ch))
