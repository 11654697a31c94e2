"""The patch-extraction oracle (oracle/warp.py) against OpenCV itself -- the third-party dependency in which the
reference's ``perspective_crop`` arithmetic lives (vae-gan.py:163-188; opencv-python, requirements.txt:4).  Bit-exact."""
import os
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import warp  # noqa: E402


def cv_crop(img, bbox, out_shape):
    w, h = out_shape
    src = np.array(bbox, dtype=np.float32).reshape(4, 2)
    dst = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.float32)
    m = cv2.getPerspectiveTransform(src, dst)
    return cv2.warpPerspective(img, m, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)


def quads(rng, h, w, n):
    for i in range(n):
        cx, cy = rng.uniform(0.2 * w, 0.8 * w), rng.uniform(0.2 * h, 0.8 * h)
        bw, bh = rng.uniform(0.1 * w, 0.6 * w), rng.uniform(0.05 * h, 0.4 * h)
        base = np.array([[cx - bw, cy - bh], [cx + bw, cy - bh], [cx + bw, cy + bh], [cx - bw, cy + bh]])
        jitter = rng.normal(0, 0.06 * min(bw, bh) * (1 + i % 4), (4, 2))
        ang = rng.uniform(-0.5, 0.5)
        rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        yield ((base - [cx, cy]) @ rot.T + [cx, cy] + jitter).tolist()


@pytest.mark.parametrize("channels", [3, 1])
@pytest.mark.parametrize("out_shape", [(448, 64), (128, 128), (64, 32), (33, 7)])
def test_crop_is_bit_exact_against_cv2(channels, out_shape):
    rng = np.random.default_rng(7 + channels)
    h, w = 173, 301
    img = rng.integers(0, 256, size=(h, w, channels) if channels == 3 else (h, w), dtype=np.uint8)
    for bbox in quads(rng, h, w, 12):                 # several of these leave the image: BORDER_REPLICATE
        want = cv_crop(img, bbox, out_shape)
        got = warp.perspective_crop(img, bbox, out_shape)
        assert got.shape == want.shape and got.dtype == np.uint8
        assert np.array_equal(got, want), int(np.abs(got.astype(int) - want.astype(int)).max())


def test_matrix_and_axis_aligned_cases():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(80, 120, 3), dtype=np.uint8)
    # axis-aligned integer box of the output size: the crop is a plain copy
    box = [[10, 20], [10 + 63, 20], [10 + 63, 20 + 31], [10, 20 + 31]]
    assert np.array_equal(warp.perspective_crop(img, box, (64, 32)), img[20:52, 10:74])
    for bbox in quads(rng, 80, 120, 8):
        src = np.array(bbox, dtype=np.float32).reshape(4, 2)
        dst = np.array([[0, 0], [447, 0], [447, 63], [0, 63]], dtype=np.float32)
        m = cv2.getPerspectiveTransform(src, dst)
        assert np.array_equal(warp.crop_matrix(bbox, (448, 64)), m)                   # bit for bit
        assert np.array_equal(warp.inverse_map(m), cv2.invert(m)[1])
    # a quadrilateral far outside the image replicates the border pixel
    far = [[500, 500], [600, 500], [600, 520], [500, 520]]
    out = warp.perspective_crop(img, far, (32, 8))
    assert np.array_equal(out, cv_crop(img, far, (32, 8))) and (out == img[-1, -1]).all()


def test_to_tensor_matches_torchvision():
    import torch
    T = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    rng = np.random.default_rng(5)
    patch = rng.integers(0, 256, size=(16, 24, 3), dtype=np.uint8)
    want = T.ToTensor()(Image.fromarray(patch)).numpy()
    assert np.array_equal(warp.to_tensor(patch), want)
    mask = rng.integers(0, 256, size=(16, 24), dtype=np.uint8)
    assert np.array_equal(warp.to_tensor(mask), T.ToTensor()(Image.fromarray(mask)).numpy())
    assert torch.from_numpy(warp.to_tensor(patch)).dtype == torch.float32


def test_oracle_reproduces_the_reference_fixture():
    """tests/golden/warp_crop.npz holds what the REFERENCE's own perspective_crop + T.ToTensor() returned for the
    deterministic inputs of oracle.warp.fixture_inputs() (PIL in, PIL out, vae-gan.py:163-188, 275-281)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_crop.npz"))
    page, mask, boxes = warp.fixture_inputs()
    assert len(gold.files) == 24
    for shape in ((448, 64), (64, 32)):
        for i, box in enumerate(boxes):
            rgb = gold[f"{shape[0]}x{shape[1]}_{i}_rgb"]            # uint8 view of the ToTensor() output, (3, H, W)
            assert np.array_equal(warp.perspective_crop(page, box, shape).transpose(2, 0, 1), rgb)
            assert np.array_equal(warp.to_tensor(warp.perspective_crop(page, box, shape)), rgb.astype(np.float32) / np.float32(255))
            assert np.array_equal(warp.perspective_crop(mask, box, shape)[None], gold[f"{shape[0]}x{shape[1]}_{i}_mask"])


def test_unwarp_is_bit_exact_against_cv2():
    """perspective_unwarp (vae-gan.py:190-200), statement for statement with cv2 itself: getPerspectiveTransform(patch
    rectangle -> bbox) + warpPerspective(dst=zero canvas, INTER_LINEAR, BORDER_TRANSPARENT).  Pins which destination
    pixels BORDER_TRANSPARENT writes in the opencv-python of this image (4.13) and their values."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for patch, bbox, cshape in warp.unwarp_cases(rng, 90):
        h, w = patch.shape[:2]
        m = cv2.getPerspectiveTransform(np.float32([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]]), np.float32(bbox).reshape(4, 2))
        canvas = np.zeros(cshape, dtype=np.uint8)
        cv2.warpPerspective(patch, m, (cshape[1], cshape[0]), dst=canvas, borderMode=cv2.BORDER_TRANSPARENT, flags=cv2.INTER_LINEAR)
        got = warp.perspective_unwarp(patch, bbox, cshape)
        assert np.array_equal(got, canvas), (patch.shape, cshape, int((got != canvas).sum()))
        assert np.array_equal(warp.unwarp_matrix(bbox, (w, h)), m)


def test_oracle_reproduces_the_reference_unwarp_fixture():
    """tests/golden/warp_unwarp.npz: canvases returned by the reference's own ``perspective_unwarp`` (vae-gan.py:190-200)."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_unwarp.npz"))
    page, mask, boxes = warp.fixture_inputs()
    n = 0
    for name, patch in warp.unwarp_fixture_patches().items():
        for i, box in enumerate(boxes):
            shape = page.shape if patch.ndim == 3 else mask.shape
            assert np.array_equal(warp.perspective_unwarp(patch, box, shape), gold[f"{name}_{i}"]), (name, i)
            n += 1
    assert n == len(gold.files) == 12
