"""GPU patch extraction (vae_gan_mark_b200/data.py -> vg_warp_perspective_u8) against the oracle of the reference's
``perspective_crop`` + ``T.ToTensor()`` (oracle/warp.py, pinned to cv2 in tests/test_warp_oracle.py).  Byte work:
the uint8 patches and the float tensors must be identical, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import warp

pytestmark = pytest.mark.gpu


def quads(rng, h, w, n):
    for i in range(n):
        cx, cy = rng.uniform(0.2 * w, 0.8 * w), rng.uniform(0.2 * h, 0.8 * h)
        bw, bh = rng.uniform(0.1 * w, 0.6 * w), rng.uniform(0.05 * h, 0.4 * h)
        base = np.array([[cx - bw, cy - bh], [cx + bw, cy - bh], [cx + bw, cy + bh], [cx - bw, cy + bh]])
        ang = rng.uniform(-0.5, 0.5)
        rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        yield ((base - [cx, cy]) @ rot.T + [cx, cy] + rng.normal(0, 0.05 * min(bw, bh) * (1 + i % 4), (4, 2))).tolist()


@pytest.mark.parametrize("channels", [3, 1])
@pytest.mark.parametrize("out_shape", [(448, 64), (128, 128), (256, 256), (33, 7)])
def test_perspective_crop_is_bit_exact(channels, out_shape):
    from vae_gan_mark_b200 import data
    rng = np.random.default_rng(31 + channels)
    h, w = 311, 517
    img = rng.integers(0, 256, size=(h, w, channels) if channels == 3 else (h, w), dtype=np.uint8)
    dev = torch.from_numpy(img).cuda()
    for bbox in quads(rng, h, w, 8):                  # several leave the image: BORDER_REPLICATE
        want = warp.perspective_crop(img, bbox, out_shape)
        got_u8 = data.perspective_crop(dev, bbox, out_shape, to_tensor=False).cpu().numpy()
        got_f = data.perspective_crop(dev, bbox, out_shape).cpu().numpy()
        assert got_u8.shape == want.shape and np.array_equal(got_u8, want)
        assert got_f.dtype == np.float32 and np.array_equal(got_f, warp.to_tensor(want))


def test_crop_batch_and_strided_source():
    from vae_gan_mark_b200 import data
    rng = np.random.default_rng(40)
    pages = [rng.integers(0, 256, size=(120 + 10 * i, 200 + 7 * i, 3), dtype=np.uint8) for i in range(4)]
    boxes = [next(quads(rng, p.shape[0], p.shape[1], 1)) for p in pages]
    batch = data.crop_batch([torch.from_numpy(p).cuda() for p in pages], boxes, (448, 64))
    assert tuple(batch.shape) == (4, 3, 64, 448) and batch.dtype == torch.float32
    for i in range(4):
        assert np.array_equal(batch[i].cpu().numpy(), warp.to_tensor(warp.perspective_crop(pages[i], boxes[i], (448, 64))))
    # a view into a wider buffer (row stride > W * C) is read in place
    wide = torch.from_numpy(rng.integers(0, 256, size=(90, 160, 3), dtype=np.uint8)).cuda()
    view = wide[:, 20:140]
    box = next(quads(rng, 90, 120, 1))
    got = data.perspective_crop(view, box, (64, 32), to_tensor=False).cpu().numpy()
    assert np.array_equal(got, warp.perspective_crop(view.cpu().numpy().copy(), box, (64, 32)))
    with pytest.raises(RuntimeError):
        data.perspective_crop(wide.cpu(), box, (64, 32))


def test_gpu_reproduces_the_reference_fixture():
    """The CUDA path against the outputs of the reference's own perspective_crop + T.ToTensor() (tests/golden/warp_crop.npz)."""
    import os
    from vae_gan_mark_b200 import data
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_crop.npz"))
    page, mask, boxes = warp.fixture_inputs()
    dpage, dmask = torch.from_numpy(page).cuda(), torch.from_numpy(mask).cuda()
    for shape in ((448, 64), (64, 32)):
        for i, box in enumerate(boxes):
            got = data.perspective_crop(dpage, box, shape).cpu().numpy()
            assert np.array_equal(got, gold[f"{shape[0]}x{shape[1]}_{i}_rgb"].astype(np.float32) / np.float32(255))
            gotm = data.perspective_crop(dmask, box, shape, to_tensor=False).cpu().numpy()
            assert np.array_equal(gotm[None], gold[f"{shape[0]}x{shape[1]}_{i}_mask"])


def test_perspective_unwarp_is_bit_exact():
    """The inverse direction (vae-gan.py:190-200) on the device against the cv2-pinned oracle: BORDER_TRANSPARENT paste of
    1-, 3- and 4-channel patches into a zero canvas and into an existing page."""
    from vae_gan_mark_b200 import data
    rng = np.random.default_rng(8)
    n = 0
    for k, (patch, bbox, cshape) in enumerate(warp.unwarp_cases(rng, 48)):
        if patch.ndim == 3 and patch.shape[2] == 2:
            continue
        h, w = patch.shape[:2]
        got = data.perspective_unwarp(torch.from_numpy(patch).cuda(), bbox.tolist(), cshape).cpu().numpy()
        assert np.array_equal(got, warp.perspective_unwarp(patch, bbox, cshape)), k
        if k % 3 == 0:
            page = rng.integers(0, 256, size=cshape, dtype=np.uint8)
            want = warp.warp_perspective_u8(patch, warp.unwarp_matrix(bbox, (w, h)), (cshape[1], cshape[0]), transparent_into=page)
            got = data.perspective_unwarp(torch.from_numpy(patch).cuda(), bbox.tolist(), cshape,
                                          canvas=torch.from_numpy(page.copy()).cuda()).cpu().numpy()
            assert np.array_equal(got, want)
        n += 1
    assert n >= 30
    # crop -> unwarp round trip at the reference's patch shape: inside the quadrilateral the page comes back within the
    # two bilinear resamplings' smoothing, outside it the canvas stays zero
    page = np.kron(rng.integers(0, 256, size=(30, 50, 3), dtype=np.uint8), np.ones((8, 8, 1), dtype=np.uint8))      # 240 x 400, blocky
    box = [[60.0, 50.0], [340.0, 60.0], [330.0, 180.0], [70.0, 170.0]]
    dev = torch.from_numpy(page).cuda()
    patch = data.perspective_crop(dev, box, (448, 64), to_tensor=False)
    back = data.perspective_unwarp(patch, box, page.shape).cpu().numpy()
    inside = back.any(axis=2)
    assert 0.2 < inside.mean() < 0.5
    diff = np.abs(back.astype(int) - page.astype(int))[inside]
    assert np.median(diff) <= 2, float(np.median(diff))


def test_crop_batch_is_one_launch():
    from vae_gan_mark_b200 import _lib, data
    rng = np.random.default_rng(41)
    pages = [torch.from_numpy(rng.integers(0, 256, size=(100 + 3 * i, 180 + 5 * i, 3), dtype=np.uint8)).cuda() for i in range(16)]
    boxes = [next(quads(rng, p.shape[0], p.shape[1], 1)) for p in pages]
    n0 = _lib.lib().vg_launch_count()
    batch = data.crop_batch(pages, boxes, (128, 128))
    assert _lib.lib().vg_launch_count() - n0 == 1
    for i in range(16):
        assert np.array_equal(batch[i].cpu().numpy(), warp.to_tensor(warp.perspective_crop(pages[i].cpu().numpy(), boxes[i], (128, 128))))
    # gray masks batch, and the C ABI rejects a bad job instead of launching
    masks = [p[:, :, 0].contiguous() for p in pages[:3]]
    mb = data.crop_batch(masks, boxes[:3], (64, 32))
    assert tuple(mb.shape) == (3, 1, 32, 64)
    for i in range(3):
        assert np.array_equal(mb[i].cpu().numpy(), warp.to_tensor(warp.perspective_crop(masks[i].cpu().numpy(), boxes[i], (64, 32))))


def test_gpu_reproduces_the_reference_unwarp_fixture():
    """The CUDA path against canvases returned by the reference's own perspective_unwarp (tests/golden/warp_unwarp.npz)."""
    import os
    from vae_gan_mark_b200 import data
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_unwarp.npz"))
    page, mask, boxes = warp.fixture_inputs()
    for name, patch in warp.unwarp_fixture_patches().items():
        for i, box in enumerate(boxes):
            shape = page.shape if patch.ndim == 3 else mask.shape
            got = data.perspective_unwarp(torch.from_numpy(patch).cuda(), box, shape).cpu().numpy()
            assert np.array_equal(got, gold[f"{name}_{i}"]), (name, i)
