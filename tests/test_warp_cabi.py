"""Host half of the patch extraction (vg_perspective_crop_matrix in libvaegan_b200.so: cv2.getPerspectiveTransform +
cv2.invert restated in C++) against the oracle and against OpenCV, bit for bit.  Runs without a GPU."""
import numpy as np
import pytest

from oracle import warp


def quads(rng, n):
    for _ in range(n):
        base = np.array([[40, 30], [260, 40], [270, 120], [35, 130]], dtype=np.float64)
        yield (base + rng.normal(0, 14, (4, 2))).astype(np.float32).tolist()


@pytest.mark.parametrize("out_shape", [(448, 64), (128, 128), (33, 7)])
def test_crop_matrix_is_bit_exact(out_shape):
    from vae_gan_mark_b200 import data
    rng = np.random.default_rng(11)
    for bbox in quads(rng, 100):
        got = np.array(list(data.perspective_crop_matrix(bbox, out_shape)), dtype=np.float64).reshape(3, 3)
        want = warp.inverse_map(warp.crop_matrix(bbox, out_shape))
        assert np.array_equal(got, want), (bbox, got - want)


def test_crop_matrix_against_cv2_and_error_code():
    cv2 = pytest.importorskip("cv2")
    from vae_gan_mark_b200 import data
    from vae_gan_mark_b200._lib import VgError
    rng = np.random.default_rng(12)
    w, h = 448, 64
    dst = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.float32)
    for bbox in quads(rng, 50):
        m = cv2.getPerspectiveTransform(np.array(bbox, dtype=np.float32), dst)
        want = cv2.invert(m)[1]
        got = np.array(list(data.perspective_crop_matrix(bbox, (w, h)))).reshape(3, 3)
        assert np.array_equal(got, want)
    with pytest.raises(VgError):            # four collinear points: no perspective transform exists
        data.perspective_crop_matrix([[0, 0], [1, 1], [2, 2], [3, 3]], (w, h))
    with pytest.raises(ValueError):
        data.perspective_crop_matrix([[0, 0], [1, 1], [2, 2]], (w, h))


@pytest.mark.parametrize("channels", [3, 1])
@pytest.mark.parametrize("out_shape", [(448, 64), (128, 128), (33, 7)])
def test_kernel_per_pixel_code_on_the_host_is_bit_exact(channels, out_shape):
    """vg_debug_warp_perspective_host runs the SAME per-pixel function the CUDA kernel runs (vg_warp.cu: warp_pixel) on
    host buffers: uint8 patch and float tensor must equal the oracle's (which is pinned to cv2) bit for bit, including
    quadrilaterals that leave the image (BORDER_REPLICATE) and a padded row stride."""
    import ctypes as C
    from vae_gan_mark_b200 import _lib, data
    rng = np.random.default_rng(20 + channels)
    h, w = 97, 203
    pitch = w * channels + 5                                        # rows are not densely packed
    buf = rng.integers(0, 256, size=(h, pitch), dtype=np.uint8)
    img = buf[:, :w * channels].reshape(h, w, channels)
    ow, oh = out_shape
    for k, bbox in enumerate(quads(rng, 10)):
        if k % 3 == 0:
            bbox = (np.array(bbox) * 1.6 - 60).tolist()             # partly outside the image
        minv = data.perspective_crop_matrix(bbox, out_shape)
        u8 = np.zeros((oh, ow, channels), dtype=np.uint8)
        chw = np.zeros((channels, oh, ow), dtype=np.float32)
        _lib.call("vg_debug_warp_perspective_host", buf.ctypes.data_as(C.c_void_p), h, w, channels, C.c_longlong(pitch), minv,
                  oh, ow, u8.ctypes.data_as(C.c_void_p), chw.ctypes.data_as(C.c_void_p))
        want = warp.perspective_crop(img if channels == 3 else img[:, :, 0], bbox, out_shape)
        want3 = want if want.ndim == 3 else want[:, :, None]
        assert np.array_equal(u8, want3), int(np.abs(u8.astype(int) - want3.astype(int)).max())
        assert np.array_equal(chw, warp.to_tensor(want))


def test_python_marshalling_through_the_host_twin():
    """data._crop builds the argument list of vg_warp_perspective_u8; with ``host_twin`` the identical list (minus the
    stream) goes to the host twin, so shapes, strides, channel handling and the float layout of the Python wrapper are
    checked here on CPU tensors exactly as tests/test_warp_gpu.py checks them on the device."""
    import torch
    from vae_gan_mark_b200 import data
    rng = np.random.default_rng(33)
    page = rng.integers(0, 256, size=(90, 160, 3), dtype=np.uint8)
    gray = rng.integers(0, 256, size=(90, 160), dtype=np.uint8)
    for k, bbox in enumerate(quads(rng, 6)):
        bbox = (np.array(bbox) * 0.5).tolist()
        want = warp.perspective_crop(page, bbox, (64, 32))
        assert np.array_equal(data._crop(torch.from_numpy(page), bbox, (64, 32), False, host_twin=True).numpy(), want)
        assert np.array_equal(data._crop(torch.from_numpy(page), bbox, (64, 32), True, host_twin=True).numpy(), warp.to_tensor(want))
        wantg = warp.perspective_crop(gray, bbox, (48, 16))
        gotg = data._crop(torch.from_numpy(gray), bbox, (48, 16), False, host_twin=True)
        assert tuple(gotg.shape) == (16, 48) and np.array_equal(gotg.numpy(), wantg)
        assert np.array_equal(data._crop(torch.from_numpy(gray), bbox, (48, 16), True, host_twin=True).numpy(), warp.to_tensor(wantg))
        view = torch.from_numpy(page)[:, 20:140]                               # row stride > W * C: read in place
        assert np.array_equal(data._crop(view, bbox, (64, 32), False, host_twin=True).numpy(),
                              warp.perspective_crop(np.ascontiguousarray(page[:, 20:140]), bbox, (64, 32)))
        chan_first = torch.from_numpy(np.ascontiguousarray(page.transpose(2, 0, 1))).permute(1, 2, 0)   # not HWC-dense: copied
        assert np.array_equal(data._crop(chan_first, bbox, (64, 32), False, host_twin=True).numpy(), want)
    with pytest.raises(RuntimeError):
        data.perspective_crop(torch.from_numpy(page), quads(rng, 1).__next__(), (64, 32))     # public entry: CUDA only


def test_full_size_page_properties():
    """A page-sized image (1500 x 2100 RGB, the scale of the reference's scans) and the reference's patch (448 x 64):
    size-independent properties -- an axis-aligned box of exactly the patch size is a plain copy, a box flipped left-right
    mirrors it -- plus oracle equality on boxes that span and that leave the page."""
    import ctypes as C
    from vae_gan_mark_b200 import _lib, data
    rng = np.random.default_rng(50)
    h, w = 1500, 2100
    page = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)

    def run(bbox, out_shape=(448, 64)):
        ow, oh = out_shape
        u8 = np.zeros((oh, ow, 3), dtype=np.uint8)
        _lib.call("vg_debug_warp_perspective_host", page.ctypes.data_as(C.c_void_p), h, w, 3, C.c_longlong(w * 3),
                  data.perspective_crop_matrix(bbox, out_shape), oh, ow, u8.ctypes.data_as(C.c_void_p), None)
        return u8

    x0, y0 = 801, 1203
    box = [[x0, y0], [x0 + 447, y0], [x0 + 447, y0 + 63], [x0, y0 + 63]]
    assert np.array_equal(run(box), page[y0:y0 + 64, x0:x0 + 448])
    flipped = [box[1], box[0], box[3], box[2]]
    assert np.array_equal(run(flipped), page[y0:y0 + 64, x0:x0 + 448][:, ::-1])
    for bbox in ([[100.3, 50.7], [1900.2, 80.1], [1880.9, 400.4], [90.6, 380.2]], [[-50, -20], [600, 10], [580, 200], [-40, 180]]):
        assert np.array_equal(run(bbox), warp.perspective_crop(page, bbox, (448, 64)))


def test_host_twin_reproduces_the_reference_fixture():
    """The library's per-pixel code against the outputs of the reference's own perspective_crop (tests/golden/warp_crop.npz)."""
    import os
    import torch
    from vae_gan_mark_b200 import data
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_crop.npz"))
    page, mask, boxes = warp.fixture_inputs()
    for shape in ((448, 64), (64, 32)):
        for i, box in enumerate(boxes):
            got = data._crop(torch.from_numpy(page), box, shape, True, host_twin=True).numpy()
            assert np.array_equal(got, gold[f"{shape[0]}x{shape[1]}_{i}_rgb"].astype(np.float32) / np.float32(255))
            gotm = data._crop(torch.from_numpy(mask), box, shape, False, host_twin=True).numpy()
            assert np.array_equal(gotm[None], gold[f"{shape[0]}x{shape[1]}_{i}_mask"])


def test_unwarp_host_half_and_per_pixel_code_are_bit_exact():
    """perspective_unwarp (vae-gan.py:190-200): vg_perspective_unwarp_matrix against the oracle's matrix, and the kernel's
    per-pixel code in BORDER_TRANSPARENT mode (host twin, through data.perspective_unwarp's own marshalling) against the
    oracle, which tests/test_warp_oracle.py pins to cv2 -- incl. pasting into a non-zero canvas."""
    import torch
    from vae_gan_mark_b200 import data
    rng = np.random.default_rng(6)
    for k, (patch, bbox, cshape) in enumerate(warp.unwarp_cases(rng, 36)):
        h, w = patch.shape[:2]
        got_m = np.array(list(data.perspective_unwarp_matrix(bbox.tolist(), (w, h)))).reshape(3, 3)
        assert np.array_equal(got_m, warp.inverse_map(warp.unwarp_matrix(bbox, (w, h))))
        if patch.ndim == 3 and patch.shape[2] == 2:
            continue                                  # the oracle handles any channel count; keep the twin to 1 / 3 / 4
        got = data.perspective_unwarp(torch.from_numpy(patch), bbox.tolist(), cshape, host_twin=True).numpy()
        assert np.array_equal(got, warp.perspective_unwarp(patch, bbox, cshape)), k
        if k % 4 == 0:
            page = rng.integers(0, 256, size=cshape, dtype=np.uint8)
            want = warp.warp_perspective_u8(patch, warp.unwarp_matrix(bbox, (w, h)), (cshape[1], cshape[0]), transparent_into=page)
            got = data.perspective_unwarp(torch.from_numpy(patch), bbox.tolist(), cshape, canvas=torch.from_numpy(page.copy()),
                                          host_twin=True).numpy()
            assert np.array_equal(got, want)
