"""BatchNorm batch statistics from the epilogue of the producing tensor-core kernel (VgConvFprop.stats), and the
deterministic (two-stage) split reduction of the weight-gradient kernel (VgConvWgrad.workspace).

Reference pattern: Conv -> BN -> ReLU everywhere in the generators (vae-gan-v2.py:172-177,237-241; vae-gan.py:52-55,
77-80).  The fused statistics must equal what the separate statistics kernel computes from the stored tensor (same
values, different summation order: relative 1e-5 on the sums), for every launch variant that feeds a BatchNorm:
plain and wide (256-column) tiles, the halo mode of narrow 3x3 layers, image-side im2col convs, stride-2 convs, and the
three ConvTranspose2d variants (pixel shuffle, column kernel, parity groups of the 4x4 stride-2 layers).
"""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def act(n, h, w, c, seed):
    from vae_gan_mark_b200.conv import new_act
    g = torch.Generator().manual_seed(seed)
    t = new_act(n, h, w, c, "cuda")
    t.copy_(torch.randn(n, h, w, c, generator=g).to(torch.bfloat16))
    return t


CONVS = [  # cin, cout, k, stride, pad, n, h, w, bias
    (64, 64, 3, 1, 1, 8, 128, 128, False),      # halo mode (>= 65 536 pixels, 64 output channels), resident weights
    (128, 64, 3, 1, 1, 4, 128, 128, False),     # halo mode, weights streamed
    (512, 512, 3, 1, 1, 8, 32, 32, False),      # 256-wide tiles, two N tiles
    (256, 1024, 3, 1, 1, 3, 8, 8, True),        # four N tiles, bias, a partly filled pixel tile
    (128, 256, 3, 2, 1, 5, 16, 16, True),       # stride 2 (vae-gan.py encoder)
    (64, 128, 3, 1, 1, 2, 20, 12, False),       # ragged pixel tiles (masked rows)
    (32, 32, 3, 1, 1, 2, 16, 16, False),        # 32 channels in 64-wide storage (vae-gan-oldv.py)
]


@pytest.mark.parametrize("cin,cout,k,s,p,n,h,w,bias", CONVS)
def test_conv_epilogue_statistics(cin, cout, k, s, p, n, h, w, bias):
    from vae_gan_mark_b200 import layers as L, ops
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(3)
    conv = nn.Conv2d(cin, cout, k, s, p, bias=bias).cuda()
    x = act(n, h, w, cin, 11)
    stats = torch.full((1, 2, cout), 7.0, device="cuda")           # the call must zero it
    op = ConvLinear(cin, cout, k, k, s, (p, p))
    with torch.no_grad():
        y = L.Conv2dFn.apply(x, conv.weight, conv.bias, op, L.WeightCache(), 0, None, None, None, stats)
    want = ops.norm_stats(y, per_sample=False)
    torch.cuda.synchronize()
    e0, e1 = rel(stats[0, 0], want[0, 0]), rel(stats[0, 1], want[0, 1])
    print(f"sum {e0:.2e}  sum of squares {e1:.2e}")
    assert e0 < 1e-4 and e1 < 1e-5, (e0, e1)
    # and the result itself is untouched by the statistics path
    with torch.no_grad():
        y2 = L.Conv2dFn.apply(x, conv.weight, conv.bias, op, L.WeightCache(), 0, None, None, None, None)
    assert torch.equal(y, y2)


CONVTS = [  # cin, cout, kh, kw, stride, pad, n, h, w   (ConvTranspose2d(cin -> cout); input h x w)
    (128, 64, 2, 2, 2, 0, 4, 16, 16),        # pixel shuffle (U-Net up-convolutions)
    (1024, 512, 4, 4, 2, 1, 6, 4, 4),        # parity groups (vae-gan.py decoder)
    (128, 64, 4, 4, 2, 1, 3, 16, 8),
    (640, 1024, 8, 1, 1, 0, 5, 1, 8),        # column kernel (bottleneck_proc of vae-gan-v2.py)
    (192, 1024, 4, 4, 1, 0, 16, 1, 1),       # full kernel from a 1x1 input (vae-gan.py decode.0)
]


@pytest.mark.parametrize("cin,cout,kh,kw,s,p,n,h,w", CONVTS)
def test_conv_transpose_epilogue_statistics(cin, cout, kh, kw, s, p, n, h, w):
    from vae_gan_mark_b200 import layers as L, ops
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(4)
    ct = nn.ConvTranspose2d(cin, cout, (kh, kw), s, p).cuda()
    oh, ow = (h - 1) * s - 2 * p + kh, (w - 1) * s - 2 * p + kw
    x = act(n, h, w, cin, 12)
    stats = torch.full((1, 2, cout), -3.0, device="cuda")
    op = ConvLinear(cout, cin, kh, kw, s, (p, p), (oh, ow))
    with torch.no_grad():
        y = L.ConvTranspose2dFn.apply(x, ct.weight, ct.bias, op, L.WeightCache(), 0, None, (oh, ow), stats)
    assert tuple(y.shape) == (n, oh, ow, cout)
    want = ops.norm_stats(y, per_sample=False)
    torch.cuda.synchronize()
    e0, e1 = rel(stats[0, 0], want[0, 0]), rel(stats[0, 1], want[0, 1])
    print(f"sum {e0:.2e}  sum of squares {e1:.2e}")
    assert e0 < 1e-4 and e1 < 1e-5, (e0, e1)


def test_image_conv_epilogue_statistics():
    from vae_gan_mark_b200 import layers as L, ops
    torch.manual_seed(5)
    conv = nn.Conv2d(4, 64, 3, 1, 1, bias=False).cuda()
    img, mask = torch.rand(6, 3, 64, 48, device="cuda"), (torch.rand(6, 1, 64, 48, device="cuda") > 0.5).float()
    stats = torch.empty((1, 2, 64), device="cuda")
    with torch.no_grad():
        y = L.ImageConvFn.apply(conv.weight, None, (3, 3, 1, 1), L.WeightCache(), 0, None, stats, img, mask)
    want = ops.norm_stats(y, per_sample=False)
    assert rel(stats[0, 0], want[0, 0]) < 1e-4 and rel(stats[0, 1], want[0, 1]) < 1e-5


@pytest.mark.parametrize("pool", [False, True])
def test_conv_bn_relu_with_fused_statistics_matches_separate_pass(pool):
    """The module-level path (run_conv_bn_relu): forward, running statistics, dx, dW, dgamma, dbeta with the statistics
    taken from the conv epilogue == with the separate statistics kernel."""
    from vae_gan_mark_b200 import modules as M
    torch.manual_seed(6)
    res = {}
    for fused in (True, False):
        M.FUSE_BN_STATS = fused
        try:
            torch.manual_seed(6)
            conv, bn = nn.Conv2d(128, 256, 3, 1, 1, bias=False).cuda(), nn.BatchNorm2d(256).cuda().train()
            x = act(4, 32, 32, 128, 13).requires_grad_()
            y, pl = M.run_conv_bn_relu(conv, bn, x, pool=pool)
            g = torch.Generator().manual_seed(2)
            gy = torch.randn(y.shape, generator=g).to(torch.bfloat16).cuda()
            if pool:
                gp = torch.randn(pl.shape, generator=g).to(torch.bfloat16).cuda()
                torch.autograd.backward([y, pl], [gy, gp])
            else:
                y.backward(gy)
            res[fused] = dict(y=y.detach().float(), dx=x.grad.float(), dw=conv.weight.grad, dg=bn.weight.grad, db=bn.bias.grad,
                              rm=bn.running_mean.clone(), rv=bn.running_var.clone(), nbt=int(bn.num_batches_tracked))
        finally:
            M.FUSE_BN_STATS = True
    a, b = res[True], res[False]
    assert a["nbt"] == b["nbt"] == 1
    for k in ("y", "dx", "dw", "dg", "db", "rm", "rv"):
        e = rel(a[k], b[k])
        print(k, f"{e:.2e}")
        assert e < 2e-3, (k, e)          # bf16 outputs may differ by one rounding where the statistics differ in the last bit


def test_wgrad_two_stage_reduction_is_deterministic_and_equal():
    """VgConvWgrad.workspace: bit-identical results from run to run, and equal (to fp32 summation order) to the atomic path."""
    from vae_gan_mark_b200 import conv
    from vae_gan_mark_b200.conv import ConvLinear
    op = ConvLinear(128, 256, 3, 3, 1, (1, 1))
    x, dy = act(8, 64, 64, 128, 21), act(8, 64, 64, 256, 22)
    outs = {}
    for det in (True, False):
        conv.DETERMINISTIC = det
        try:
            outs[det] = [op.backward_weight(dy, x).contiguous().clone() for _ in range(3)]
        finally:
            conv.DETERMINISTIC = False
    torch.cuda.synchronize()
    assert torch.equal(outs[True][0], outs[True][1]) and torch.equal(outs[True][0], outs[True][2])
    e = rel(outs[True][0], outs[False][0])
    print("two-stage vs atomic:", f"{e:.2e}")
    assert e < 1e-5
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (256, 128, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
    assert rel(outs[True][0], ref) < 1e-3


@pytest.mark.parametrize("cin,cout,k,s,p,n,h,w", [(512, 512, 3, 1, 1, 4, 32, 32), (128, 256, 3, 1, 1, 8, 32, 32),
                                                  (512, 1024, 1, 1, 0, 8, 32, 32), (1024, 384, 3, 1, 1, 8, 16, 16),
                                                  (128, 256, 3, 2, 1, 8, 64, 64), (192, 320, 3, 1, 1, 4, 24, 40)])
def test_wgrad_cta_pair_kernel_equals_single_cta_kernel(cin, cout, k, s, p, n, h, w):
    """conv_wgrad_kernel<true> (clusters of two CTAs, tcgen05.mma.cta_group::2 with M = 256, each CTA staging half of the
    X tile) against the single-CTA kernel (same products, same fp32 accumulation, only the split of the work differs) and
    against torch: odd numbers of 128-channel tiles (cout 384 / 320: the pair's second CTA idles or is partly masked),
    partial N tiles, stride 2, 1x1."""
    from vae_gan_mark_b200 import _lib
    from vae_gan_mark_b200.conv import ConvLinear
    op = ConvLinear(cin, cout, k, k, s, (p, p))
    x = act(n, h, w, cin, 31)
    oh, ow = op.out_hw(h, w)
    dy = act(n, oh, ow, cout, 32)
    res = {}
    try:
        for pairs in (0, 1):
            _lib.lib().vg_set_cta_pairs(pairs)
            res[pairs] = op.backward_weight(dy, x).contiguous().clone()
    finally:
        _lib.lib().vg_set_cta_pairs(1)
    torch.cuda.synchronize()
    e = rel(res[1], res[0])
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, k, k), dy.float().permute(0, 3, 1, 2),
                                      stride=s, padding=p)
    print(f"pair vs single {e:.2e}, pair vs torch {rel(res[1], ref):.2e}")
    assert e < 1e-5 and rel(res[1], ref) < 2e-3


@pytest.mark.parametrize("cin,cout,k,s,p,n,h,w,bias", [(512, 512, 3, 1, 1, 8, 64, 64, False), (256, 512, 3, 1, 1, 37, 16, 16, True),
                                                       (512, 1024, 1, 1, 0, 16, 32, 32, True), (128, 256, 3, 2, 1, 64, 64, 64, True),
                                                       (512, 384, 3, 1, 1, 33, 24, 20, False),
                                                       (128, 128, 3, 1, 1, 16, 64, 64, True),       # 128-wide N tiles in pairs
                                                       (128, 512, 3, 1, 1, 8, 64, 64, False)])      # ... for the data gradient
def test_fprop_cta_pair_kernel_is_bit_equal_to_single_cta_kernel(cin, cout, k, s, p, n, h, w, bias):
    """conv_fprop_kernel<true> (clusters of two CTAs, one stream of M = 256 tcgen05.mma.cta_group::2, each CTA staging its own
    128-pixel box and half of the weight tile) must reproduce the single-CTA kernel bit for bit -- same products, same
    K order, same fp32 accumulator: forward (+ fused ReLU, bias, BatchNorm statistics), data gradient, and the pixel-shuffle
    epilogue of a 2x2 stride-2 transposed conv; odd pixel-tile counts (n = 37 / 33) leave the last pair half empty."""
    from vae_gan_mark_b200 import _lib, layers as L
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(8)
    op = ConvLinear(cin, cout, k, k, s, (p, p))
    x = act(n, h, w, cin, 41)
    wt = (torch.randn(cout, cin, k, k) * (cin * k * k) ** -0.5).cuda()
    b = torch.randn(cout).cuda() if bias else None
    oh, ow = op.out_hw(h, w)
    dy = act(n, oh, ow, cout, 42)
    wf, wb = op.prep_fwd(wt), op.prep_bwd(wt)
    ct = nn.ConvTranspose2d(cout, 128, 2, 2).cuda()
    opt = ConvLinear(128, cout, 2, 2, 2, (0, 0), (2 * oh, 2 * ow))
    res = {}
    try:
        for pairs in (0, 1):
            _lib.lib().vg_set_fprop_cta_pairs(pairs)
            stats = torch.empty((1, 2, cout), device="cuda")
            y = op.forward(x, wf, b, 1, stats=stats)
            dx = op.backward_data(dy, wb, (h, w))
            dx_mn = op.backward_data(dy, {"mn": wf}, (h, w))      # the forward operand read MN-major (pairs: half the boxes per CTA)
            with torch.no_grad():
                up = L.ConvTranspose2dFn.apply(dy, ct.weight, ct.bias, opt, L.WeightCache(), 0, None, (2 * oh, 2 * ow))
            res[pairs] = (y.clone(), dx.clone(), up.clone(), stats.clone(), dx_mn.clone())
    finally:
        _lib.lib().vg_set_fprop_cta_pairs(1)
    torch.cuda.synchronize()
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    assert torch.equal(res[0][4], res[1][4]) and rel(res[1][4], res[1][1]) < 1e-6       # MN-major operand: pairs == single == K-major
    assert rel(res[1][3], res[0][3]) < 1e-5
    ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float(), b, stride=s, padding=p))
    assert rel(res[1][0].float().permute(0, 3, 1, 2), ref) < 1e-2


@pytest.mark.parametrize("cin,cout,n,h,w", [(64, 64, 8, 128, 128), (128, 64, 4, 64, 96), (64, 64, 3, 61, 75), (64, 32, 2, 72, 130),
                                            (192, 64, 2, 64, 64)])
@pytest.mark.parametrize("deterministic", [False, True])
def test_wgrad_dual_shift_mode_matches_torch(cin, cout, n, h, w, deterministic):
    """Weight gradient of 3x3 stride-1 layers with <= 64 output channels: the kernel fills the second half of its M = 128 rows
    with dY shifted one pixel column (six X shifts instead of nine taps, vg_conv_wgrad.cu: wgrad_plan).  Against
    torch.nn.grad.conv2d_weight on the same bf16 inputs, incl. widths / heights that are not multiples of the 8 x 8 pixel
    tile (the extra tile column that carries the shifted last column) and the deterministic two-stage reduction."""
    from vae_gan_mark_b200 import conv
    from vae_gan_mark_b200.conv import ConvLinear
    op = ConvLinear(cin, cout, 3, 3, 1, (1, 1))
    x = act(n, h, w, cin, 51)
    dy = act(n, h, w, cout, 52)
    old = conv.DETERMINISTIC
    try:
        conv.DETERMINISTIC = deterministic
        got = op.backward_weight(dy, x).contiguous().clone()
        again = op.backward_weight(dy, x).contiguous().clone()
    finally:
        conv.DETERMINISTIC = old
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3), dy[..., :cout].float().permute(0, 3, 1, 2),
                                      stride=1, padding=1)
    e = rel(got, ref)
    print(f"dual-shift wgrad {cin}->{cout} {n}x{h}x{w}: {e:.2e}")
    assert e < 2e-3
    for tap in range(9):            # every tap separately: a wrong (row half, block) -> tap map would swap whole taps
        r, q = divmod(tap, 3)
        assert rel(got[:, :, r, q], ref[:, :, r, q]) < 4e-3, (tap, rel(got[:, :, r, q], ref[:, :, r, q]))
    if deterministic:
        assert torch.equal(got, again)
