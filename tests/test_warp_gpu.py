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
