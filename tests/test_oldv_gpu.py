"""GPU parity of the vae-gan-oldv.py family's extra pieces against plain PyTorch fp32
references of the same ops on the same bf16-rounded inputs, through the C ABI:

  * the separable 2-D bilinear resize of the 4-row text map (vae-gan-oldv.py:165-176, 286-291),
  * the per-channel skip gate written into a concat slice (vae-gan-oldv.py:226-231),
  * the 32-channel levels of the 3-level U-Net on the tensor-core kernels (channel counts below one 64-wide K block),
  * the Conv1d + positional-encoding text encoder (vae-gan-oldv.py:74-148) against the oracle's module.

The full training step of the family is held to the CPU oracle in tests/test_step_parity_gpu.py (bf16) and
tests/test_fp32_mode_gpu.py (fp32 mode).  Tolerance: relative L2 <= 1e-2 per tensor in bf16 mode (see test_ops_gpu.py).
"""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def bf(t):
    return t.to(torch.bfloat16).float()


def check(name, got, want, tol=TOL):
    e = rel(got, want)
    print(f"{name}: {e:.2e}")
    assert e <= tol, (name, e)


def act_leaf(t_nchw, dtype=None):
    """NCHW fp32 (cpu) -> NHWC activation leaf on cuda, channel-padded like every activation of the package."""
    from vae_gan_mark_b200 import ops
    from vae_gan_mark_b200.conv import new_act
    n, c, h, w = t_nchw.shape
    a = new_act(n, h, w, c, "cuda", dtype)
    ops.strided_copy(t_nchw.permute(0, 2, 3, 1).cuda(), a)
    return a.detach().requires_grad_(True)


def act_grad(t_nchw, dtype=None):
    return act_leaf(t_nchw, dtype).detach()


def nchw(t):
    return t.detach().float().cpu().permute(0, 3, 1, 2)


@pytest.fixture(params=["bf16", "fp32"])
def precision(request):
    import vae_gan_mark_b200 as vg
    vg.set_precision(request.param)
    yield request.param
    vg.set_precision("bf16")


@pytest.mark.parametrize("n,c,h0,w0,h,w", [(2, 64, 4, 4, 8, 16), (3, 128, 4, 4, 1, 8), (2, 512, 4, 2, 32, 64),
                                           (1, 64, 4, 28, 4, 56), (2, 64, 4, 3, 16, 24)])
def test_upsample_2d_matches_interpolate(precision, n, c, h0, w0, h, w):
    from vae_gan_mark_b200 import layers as L
    torch.manual_seed(11)
    rnd = bf if precision == "bf16" else (lambda t: t)
    tol = TOL if precision == "bf16" else 1e-5
    t = rnd(torch.randn(n, c, h0, w0))
    rt = t.clone().requires_grad_(True)
    y_ref = F.interpolate(rt, size=(h, w), mode="bilinear", align_corners=False)
    g = rnd(torch.randn_like(y_ref))
    y_ref.backward(g)
    tc = act_leaf(t)
    y = L.Upsample2DFn.apply(tc, h, w)
    y.backward(act_grad(g))
    check("y", nchw(y), y_ref, tol)
    check("dt", nchw(tc.grad), rt.grad, tol)


@pytest.mark.parametrize("n,c,h,w", [(2, 32, 8, 12), (3, 128, 5, 7), (2, 64, 16, 16)])
def test_channel_gate_into_concat_slice(precision, n, c, h, w):
    from vae_gan_mark_b200 import layers as L
    from vae_gan_mark_b200 import modules as M
    from vae_gan_mark_b200 import ops
    torch.manual_seed(12)
    rnd = bf if precision == "bf16" else (lambda t: t)
    tol = TOL if precision == "bf16" else 1e-5
    x = rnd(torch.randn(n, c, h, w))
    alpha = (torch.randn(1, c, 1, 1) * 0.5 + 0.3)
    ra, rx = alpha.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y_ref = rx * torch.sigmoid(ra)
    g = rnd(torch.randn_like(y_ref))
    y_ref.backward(g)

    gate = M.GatedSkipConnection(c).cuda()
    with torch.no_grad():
        gate.alpha.copy_(alpha)
    xc = act_leaf(x)
    buf = torch.zeros((n, h, w, 2 * c), dtype=ops.act_dtype(), device="cuda")
    buf[..., :c] = 7.0                                          # the other half of the buffer must stay untouched
    y = gate(xc, out=buf[..., c:])
    assert y.data_ptr() == buf[..., c:].data_ptr()
    gfull = torch.zeros((n, h, w, 2 * c), dtype=ops.act_dtype(), device="cuda")
    gfull[..., c:] = g.permute(0, 2, 3, 1).cuda().to(ops.act_dtype())
    y.backward(gfull[..., c:])
    check("y", nchw(buf[..., c:]), y_ref, tol)
    assert bool((buf[..., :c] == 7.0).all())
    check("dx", nchw(xc.grad), rx.grad, tol)
    check("dalpha", gate.alpha.grad, ra.grad, 1e-4 if precision == "fp32" else TOL)


@pytest.mark.parametrize("cin,cout,h,w", [(64, 32, 16, 16), (32, 32, 16, 24), (32, 64, 8, 8), (32, 32, 256, 128)])
def test_conv3x3_with_32_channel_sides(cin, cout, h, w):
    """3x3 convs whose input and / or output has 32 channels (levels 32/64/128 of vae-gan-oldv.py:187-224): the
    activations live in 64-wide zero-padded storage, the weight operand is zero-padded along K.  The last case has
    enough pixels (>= 65536) for the halo mode of the forward kernel."""
    from vae_gan_mark_b200 import layers as L
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(13)
    n = 2
    x = bf(torch.randn(n, cin, h, w))
    conv = nn.Conv2d(cin, cout, 3, 1, 1, bias=False)
    wr = bf(conv.weight.detach()).requires_grad_(True)
    rx = x.clone().requires_grad_(True)
    y_ref = F.conv2d(rx, wr, None, 1, 1)
    g = bf(torch.randn_like(y_ref))
    y_ref.backward(g)
    conv = conv.cuda()
    xc = act_leaf(x)
    op = ConvLinear(cin, cout, 3, 3, 1, (1, 1))
    y = L.Conv2dFn.apply(xc, conv.weight, None, op, L.WeightCache(), 0, None, None, None)
    assert y.shape[3] == cout
    y.backward(act_grad(g))
    check("y", nchw(y), y_ref)
    check("dx", nchw(xc.grad), rx.grad)
    check("dw", conv.weight.grad, wr.grad)


def test_convT_32_channels_into_slice_then_small_out_conv():
    """up_tconv3 (ConvT2x2 s2 64 -> 32 into the lower half of a 64-channel concat buffer, vae-gan-oldv.py:262) and the
    32 -> 3 final 1x1 conv (vae-gan-oldv.py:267)."""
    from vae_gan_mark_b200 import layers as L
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(14)
    n, cin, c, h, w = 2, 64, 32, 8, 8
    x = bf(torch.randn(n, cin, h, w))
    ct = nn.ConvTranspose2d(cin, c, 2, 2)
    wr, br = bf(ct.weight.detach()).requires_grad_(True), ct.bias.detach().clone().requires_grad_(True)
    rx = x.clone().requires_grad_(True)
    y_ref = F.conv_transpose2d(rx, wr, br, stride=2)
    g = bf(torch.randn_like(y_ref))
    y_ref.backward(g)
    ct = ct.cuda()
    buf = torch.zeros(n, 2 * h, 2 * w, 2 * c, dtype=torch.bfloat16, device="cuda")
    buf[..., c:] = 3.0
    xc = act_leaf(x)
    op = ConvLinear(c, cin, 2, 2, 2, (0, 0), (2 * h, 2 * w))
    up = L.ConvTranspose2dFn.apply(xc, ct.weight, ct.bias, op, L.WeightCache(), 0, buf[..., :c], (2 * h, 2 * w))
    gfull = torch.randn(n, 2 * h, 2 * w, 2 * c, device="cuda").to(torch.bfloat16)      # the neighbour slice's gradient is
    gfull[..., :c] = g.permute(0, 2, 3, 1).cuda().to(torch.bfloat16)                   # non-zero, as in the real decoder
    up.backward(gfull[..., :c])
    check("y", nchw(buf[..., :c]), y_ref)
    assert bool((buf[..., c:] == 3.0).all())
    check("dx", nchw(xc.grad), rx.grad)
    check("dw", ct.weight.grad, wr.grad)
    check("db", ct.bias.grad, br.grad)

    conv = nn.Conv2d(c, 3, 1)
    x2 = bf(torch.randn(n, c, 16, 16))
    rx2 = x2.clone().requires_grad_(True)
    o_ref = torch.sigmoid(conv(rx2))
    g2 = torch.randn_like(o_ref)
    o_ref.backward(g2)
    convc = nn.Conv2d(c, 3, 1).cuda()
    convc.load_state_dict(conv.state_dict())
    x2c = act_leaf(x2)
    o = L.SigmoidOutFn.apply(L.SmallOutConvFn.apply(x2c, convc.weight, convc.bias, 0))
    o.backward(g2.cuda())
    check("out", o, o_ref, 1e-4)
    check("out dx", nchw(x2c.grad), rx2.grad)
    check("out dw", convc.weight.grad, conv.weight.grad, 1e-4)


def test_norm_act_pool_32_channels():
    from vae_gan_mark_b200 import layers as L
    torch.manual_seed(15)
    n, c, h, w = 3, 32, 8, 12
    x = bf(torch.randn(n, c, h, w) * 1.5 + 0.3)
    gamma, beta = torch.rand(c) + 0.5, torch.randn(c) * 0.2
    rx = x.clone().requires_grad_(True)
    g_ref, b_ref = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y_ref = F.relu(F.batch_norm(rx, torch.zeros(c), torch.ones(c), g_ref, b_ref, True, 0.1, 1e-5))
    p_ref = F.max_pool2d(y_ref, 2, 2)
    gy, gp = bf(torch.randn_like(y_ref)), bf(torch.randn_like(p_ref))
    ((y_ref * gy).sum() + (p_ref * gp).sum()).backward()
    xc = act_leaf(x)
    gc, bc = gamma.cuda().requires_grad_(True), beta.cuda().requires_grad_(True)
    state = {"training": True, "running_mean": torch.zeros(c, device="cuda"), "running_var": torch.ones(c, device="cuda"),
             "num_batches_tracked": torch.zeros((), dtype=torch.long, device="cuda")}
    y, pl = L.NormActFn.apply(xc, gc, bc, False, 1, True, None, 1e-5, state, None)
    assert y.stride(2) == 64 and pl.stride(2) == 64
    ((y.float() * act_grad(gy).float()).sum() + (pl.float() * act_grad(gp).float()).sum()).backward()
    check("y", nchw(y), y_ref)
    check("pool", nchw(pl), p_ref)
    check("dx", nchw(xc.grad), rx.grad, 2e-2)
    check("dgamma", gc.grad, g_ref.grad)
    check("dbeta", bc.grad, b_ref.grad)


@pytest.mark.parametrize("batch,w0", [(3, 4), (16, 28)])
def test_oldv_text_encoder_matches_oracle(precision, batch, w0):
    """Embedding -> biGRU (cluster kernel) -> Conv1d as a 1x3 tensor-core conv -> pool -> 4 rows + positional encoding,
    against the oracle's CharacterTokenEncoderOldV (pinned to the reference by the oldv golden fixture) in float64."""
    from oracle import models as om
    from vae_gan_mark_b200 import modules as M
    torch.manual_seed(16)
    ref = om.CharacterTokenEncoderOldV(om.ALPHABET_STR, 128, 256, 2, w0, 4)
    mine = M.CharacterTokenEncoderOldV(M.ALPHABET_STR, 128, 256, 2, w0, 4)
    mine.load_state_dict(ref.state_dict(), strict=True)
    ref, mine = ref.double(), mine.cuda()                       # float64 reference on the CPU
    ref.rnn.dropout = mine.rnn.dropout = 0.0
    texts = ["hello world %d" % i * (1 + i % 3) for i in range(batch)]
    y_ref = ref(texts)
    g = torch.randn_like(y_ref)
    (y_ref * g).sum().backward()
    y = mine(texts)
    assert tuple(y.shape) == (batch, 512, 4, w0) and y.dtype == torch.float32
    (y * g.float().cuda()).sum().backward()
    tol = 2e-2 if precision == "bf16" else 1e-4
    check("text map", y, y_ref, tol)
    worst = 0.0
    for (name, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, name
        e = rel(p.grad, q.grad)
        worst = max(worst, e)
        assert e <= (5e-2 if precision == "bf16" else 5e-4), (name, e)
    print("worst parameter-gradient error:", f"{worst:.2e}")


def test_oldv_model_forward_shapes_and_all_parameters_get_gradients():
    from vae_gan_mark_b200 import modules as M
    torch.manual_seed(17)
    h, w, b = 32, 64, 2
    G = M.VAEGAN_UNet_SpatialFiLM_OldV(4, 128, patch_shape=(w, h)).cuda().train()
    img, mask = torch.rand(b, 3, h, w, device="cuda"), (torch.rand(b, 1, h, w, device="cuda") > 0.5).float()
    out, mu, lv = G(img, mask, ["abc", "hello world"])
    assert tuple(out.shape) == (b, 3, h, w) and tuple(mu.shape) == (b, 128, 1, 1) and tuple(lv.shape) == (b, 128, 1, 1)
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    (out.mean() + 0.01 * G.__dict__["_last_kl"]).backward()
    missing = [n for n, p in G.named_parameters() if p.grad is None]
    assert not missing, missing
    assert all(bool(torch.isfinite(p.grad).all()) for p in G.parameters())
