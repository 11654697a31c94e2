"""Per-op parity of every autograd Function (forward, input gradient, parameter gradients) against a plain
PyTorch fp32 reference of the same op, fed the same bf16-rounded inputs.  Calls go through the C ABI.

Tolerance: relative L2 error <= 1e-2 per tensor (bf16 output rounding is 2^-9 = 2e-3 per element; fp32 results
such as weight gradients are typically < 1e-4).  Unlike the whole-step test there is no deep cancellation here,
so this is where the backward kernels are pinned tightly.
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


def nhwc(t):       # NCHW fp32 -> NHWC bf16 leaf on cuda
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda().requires_grad_(True)


def nchw(t):       # NHWC -> NCHW fp32 cpu
    return t.detach().float().cpu().permute(0, 3, 1, 2)


def check(name, got, want, tol=TOL):
    e = rel(got, want)
    print(f"{name}: {e:.2e}")
    assert e <= tol, (name, e)


@pytest.mark.parametrize("cin,cout,k,s,p,h,w,bias", [
    (64, 64, 3, 1, 1, 16, 16, False), (128, 256, 3, 1, 1, 8, 12, True), (512, 512, 3, 1, 1, 8, 8, False),
    (64, 128, 4, 2, 1, 16, 16, True), (128, 256, 3, 2, 1, 16, 16, True), (512, 1024, 1, 1, 0, 4, 4, True),
    (128, 64, 2, 2, 0, 8, 8, False), (576, 64, 3, 1, 1, 8, 8, False)])
def test_conv2d(cin, cout, k, s, p, h, w, bias):
    from vae_gan_mark_b200 import layers as L
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(0)
    n = 3
    x = bf(torch.randn(n, cin, h, w))
    conv = nn.Conv2d(cin, cout, k, s, p, bias=bias)
    ref_x = x.clone().requires_grad_(True)
    conv_ref = nn.Conv2d(cin, cout, k, s, p, bias=bias)
    conv_ref.load_state_dict(conv.state_dict())
    conv_ref.weight.data = bf(conv_ref.weight.data)
    y_ref = conv_ref(ref_x)
    gy = bf(torch.randn_like(y_ref))
    y_ref.backward(gy)

    conv = conv.cuda()
    xc = nhwc(x)
    op = ConvLinear(cin, cout, k, k, s, (p, p))
    y = L.Conv2dFn.apply(xc, conv.weight, conv.bias, op, L.WeightCache(), 0, None, 0, None)
    y.backward(nhwc(gy).detach())
    check("y", nchw(y), y_ref)
    check("dx", nchw(xc.grad), ref_x.grad)
    check("dw", conv.weight.grad, conv_ref.weight.grad)
    if bias:
        check("db", conv.bias.grad, conv_ref.bias.grad)


@pytest.mark.parametrize("cin,cout,kh,kw,s,p,h,w", [
    (128, 64, 2, 2, 2, 0, 8, 8), (1024, 512, 4, 4, 2, 1, 4, 4), (128, 64, 4, 4, 2, 1, 16, 8),
    (640, 1024, 2, 1, 1, 0, 1, 4), (192, 1024, 2, 2, 1, 0, 1, 1), (544, 1024, 4, 1, 1, 0, 1, 3),
    (640, 1024, 1, 1, 1, 0, 1, 2)])
def test_conv_transpose2d(cin, cout, kh, kw, s, p, h, w):
    from vae_gan_mark_b200 import layers as L
    from vae_gan_mark_b200.conv import ConvLinear, new_act
    from vae_gan_mark_b200 import ops
    torch.manual_seed(1)
    n = 3
    x = bf(torch.randn(n, cin, h, w))
    ct = nn.ConvTranspose2d(cin, cout, (kh, kw), s, p)
    ct_ref = nn.ConvTranspose2d(cin, cout, (kh, kw), s, p)
    ct_ref.load_state_dict(ct.state_dict())
    ct_ref.weight.data = bf(ct_ref.weight.data)
    ref_x = x.clone().requires_grad_(True)
    y_ref = ct_ref(ref_x)
    gy = bf(torch.randn_like(y_ref))
    y_ref.backward(gy)
    oh, ow = y_ref.shape[2], y_ref.shape[3]

    ct = ct.cuda()
    xa = new_act(n, h, w, cin, "cuda")                      # padded buffer when cin % 64 != 0
    ops.strided_copy(x.permute(0, 2, 3, 1).cuda(), xa)
    xc = xa.detach().requires_grad_(True)
    op = ConvLinear(cout, cin, kh, kw, s, (p, p), (oh, ow))
    y = L.ConvTranspose2dFn.apply(xc, ct.weight, ct.bias, op, L.WeightCache(), 0, None, (oh, ow))
    y.backward(nhwc(gy).detach())
    check("y", nchw(y), y_ref)
    check("dx", nchw(xc.grad), ref_x.grad)
    check("dw", ct.weight.grad, ct_ref.weight.grad)
    check("db", ct.bias.grad, ct_ref.bias.grad)


def test_convT_into_concat_slice_and_cat():
    """ConvT2x2 writes channels [0,C) of a concat buffer whose upper half already holds the skip; no copies."""
    from vae_gan_mark_b200 import layers as L
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(2)
    n, cin, c, h, w = 2, 128, 64, 4, 4
    x, skip = bf(torch.randn(n, cin, h, w)), bf(torch.randn(n, c, 2 * h, 2 * w))
    ct = nn.ConvTranspose2d(cin, c, 2, 2).cuda()
    buf = torch.zeros(n, 2 * h, 2 * w, 2 * c, dtype=torch.bfloat16, device="cuda")
    buf[..., c:] = skip.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
    skip_view = buf[..., c:].detach().requires_grad_(True)
    xc = nhwc(x)
    op = ConvLinear(c, cin, 2, 2, 2, (0, 0), (2 * h, 2 * w))
    up = L.ConvTranspose2dFn.apply(xc, ct.weight, ct.bias, op, L.WeightCache(), 0, buf[..., :c], (2 * h, 2 * w))
    cat = L.CatSlicesFn.apply(up, skip_view, buf)
    assert cat.data_ptr() == buf.data_ptr()
    ref = torch.cat([F.conv_transpose2d(x, bf(ct.weight.detach().cpu()), ct.bias.detach().cpu(), stride=2), skip], 1)
    check("cat", nchw(cat), ref)
    g = bf(torch.randn_like(ref))
    cat.backward(nhwc(g).detach())
    check("dskip", nchw(skip_view.grad), g[:, c:])


@pytest.mark.parametrize("per_sample,act,pool", [(False, 1, True), (False, 1, False), (True, 2, False)])
def test_norm_act(per_sample, act, pool):
    from vae_gan_mark_b200 import layers as L
    torch.manual_seed(3)
    n, c, h, w = 4, 128, 8, 12
    x = bf(torch.randn(n, c, h, w) * 1.5 + 0.3)
    gamma, beta = torch.rand(c) + 0.5, torch.randn(c) * 0.2
    rx = x.clone().requires_grad_(True)
    g_ref, b_ref = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = torch.zeros(c), torch.ones(c)
    if per_sample:
        y_ref = F.leaky_relu(F.instance_norm(rx, weight=g_ref, bias=b_ref, eps=1e-5), 0.2)
    else:
        y_ref = F.relu(F.batch_norm(rx, rm, rv, g_ref, b_ref, True, 0.1, 1e-5))
    p_ref = F.max_pool2d(y_ref, 2, 2) if pool else None
    gy = bf(torch.randn_like(y_ref))
    gp = bf(torch.randn_like(p_ref)) if pool else None
    (y_ref * gy).sum().backward(retain_graph=pool)
    if pool:
        (p_ref * gp).sum().backward()

    xc = nhwc(x)
    gc, bc = gamma.cuda().requires_grad_(True), beta.cuda().requires_grad_(True)
    state = None if per_sample else {"training": True, "running_mean": torch.zeros(c, device="cuda"),
                                     "running_var": torch.ones(c, device="cuda"),
                                     "num_batches_tracked": torch.zeros((), dtype=torch.long, device="cuda")}
    y, pl = L.NormActFn.apply(xc, gc, bc, per_sample, act, pool, None, 1e-5, state, None)
    loss = (y.float() * nhwc(gy).detach().float()).sum()
    if pool:
        loss = loss + (pl.float() * nhwc(gp).detach().float()).sum()
    loss.backward()
    check("y", nchw(y), y_ref)
    if pool:
        check("pool", nchw(pl), p_ref)
    check("dx", nchw(xc.grad), rx.grad, 2e-2)
    check("dgamma", gc.grad, g_ref.grad)
    check("dbeta", bc.grad, b_ref.grad)
    if state is not None:
        check("running_mean", state["running_mean"], rm, 1e-3)
        check("running_var", state["running_var"], rv, 1e-3)
        assert int(state["num_batches_tracked"]) == 1


def test_film_upsample_ztext():
    from vae_gan_mark_b200 import layers as L
    torch.manual_seed(4)
    n, c, h, w, w0, t = 2, 64, 8, 16, 2, 128
    gb, x = bf(torch.randn(n, 2 * c, h, w)), bf(torch.randn(n, c, h, w))
    rgb, rx = gb.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y_ref = rgb[:, :c] * rx + rgb[:, c:]
    g = bf(torch.randn_like(y_ref))
    y_ref.backward(g)
    gbc, xc = nhwc(gb), nhwc(x)
    y = L.FiLMFn.apply(gbc, xc)
    y.backward(nhwc(g).detach())
    check("film y", nchw(y), y_ref)
    check("film dgb", nchw(gbc.grad), rgb.grad)
    check("film dx", nchw(xc.grad), rx.grad)

    tx = bf(torch.randn(n, t, 1, w0))
    rt = tx.clone().requires_grad_(True)
    u_ref = F.interpolate(rt, size=(h, w), mode="bilinear", align_corners=False)
    gu = bf(torch.randn_like(u_ref))
    u_ref.backward(gu)
    tc = nhwc(tx)
    u = L.UpsampleWFn.apply(tc, h, w)
    u.backward(nhwc(gu).detach())
    check("upsample y", nchw(u), u_ref)
    check("upsample dt", nchw(tc.grad), rt.grad)

    for zc in (128, 32):
        z = torch.randn(n, zc)
        rz, rt2 = z.clone().requires_grad_(True), tx.clone().requires_grad_(True)
        c_ref = torch.cat([rz.view(n, zc, 1, 1).expand(-1, -1, 1, w0), rt2], 1)
        gc_ = bf(torch.randn_like(c_ref))
        c_ref.backward(gc_)
        zcu, tcu = z.cuda().requires_grad_(True), nhwc(tx)
        cc = L.ZTextCatFn.apply(zcu, tcu)
        gin = L.grad_in(nhwc(gc_).detach())
        cc.backward(gin)
        check("ztext y", nchw(cc), bf(c_ref))
        check("ztext dz", zcu.grad, rz.grad)
        check("ztext dtext", nchw(tcu.grad), rt2.grad)


def test_heads_reparam_kl():
    from vae_gan_mark_b200 import layers as L
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(5)
    n, c, h, w, z = 5, 1024, 2, 4, 128
    x = bf(torch.randn(n, c, h, w))
    mu_h, lv_h = nn.Conv2d(c, z, (h, w)), nn.Conv2d(c, z, (h, w))
    eps = torch.randn(n, z)
    rx = x.clone().requires_grad_(True)
    wm, wl = bf(mu_h.weight.detach()).requires_grad_(True), bf(lv_h.weight.detach()).requires_grad_(True)
    bm, bl = mu_h.bias.detach().clone().requires_grad_(True), lv_h.bias.detach().clone().requires_grad_(True)
    mu_ref, lv_ref = F.conv2d(rx, wm, bm), F.conv2d(rx, wl, bl)
    z_ref = mu_ref + eps.view(n, z, 1, 1) * torch.exp(0.5 * lv_ref)
    kl_ref = torch.mean(-0.5 * torch.mean(1 + lv_ref - mu_ref.pow(2) - lv_ref.exp(), dim=[1, 2, 3]))
    gz = torch.randn(n, z)
    (0.37 * kl_ref + (z_ref.view(n, z) * gz).sum()).backward()

    mu_h, lv_h = mu_h.cuda(), lv_h.cuda()
    xc = nhwc(x)
    op = ConvLinear(c, 2 * z, h, w, 1, (0, 0), (h, w))
    heads = L.HeadsFn.apply(xc, mu_h.weight, lv_h.weight, op, L.WeightCache(), L.WeightCache())
    mu, lv, zz, kl = L.ReparamKLFn.apply(heads, mu_h.bias, lv_h.bias, eps.cuda())
    (0.37 * kl + (zz * gz.cuda()).sum()).backward()
    check("mu", mu, mu_ref.view(n, z))
    check("logvar", lv, lv_ref.view(n, z))
    check("z", zz, z_ref.view(n, z))
    check("kl", kl, kl_ref)
    check("dx", nchw(xc.grad), rx.grad)
    check("dw_mu", mu_h.weight.grad, wm.grad)
    check("dw_lv", lv_h.weight.grad, wl.grad)
    check("db_mu", mu_h.bias.grad, bm.grad)
    check("db_lv", lv_h.bias.grad, bl.grad)


@pytest.mark.parametrize("k,s,p,act,cin", [(4, 2, 1, 2, 3), (3, 1, 1, 0, 4), (3, 2, 1, 0, 4)])
def test_image_conv(k, s, p, act, cin):
    from vae_gan_mark_b200 import layers as L
    torch.manual_seed(6)
    n, h, w, cout = 3, 16, 24, 64
    imgs = [torch.rand(n, 3, h, w)] + ([(torch.rand(n, 1, h, w) > 0.5).float()] if cin == 4 else [])
    conv = nn.Conv2d(cin, cout, k, s, p)
    xin = bf(torch.cat(imgs, 1)).requires_grad_(True)
    wr = bf(conv.weight.detach()).requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True)
    y_ref = F.conv2d(xin, wr, br, s, p)
    if act == 2:
        y_ref = F.leaky_relu(y_ref, 0.2)
    g = bf(torch.randn_like(y_ref))
    y_ref.backward(g)
    conv = conv.cuda()
    cimgs = [t.cuda().requires_grad_(True) for t in imgs]
    y = L.ImageConvFn.apply(conv.weight, conv.bias, (k, k, s, p), L.WeightCache(), act, None, None, *cimgs)
    y.backward(nhwc(g).detach())
    check("y", nchw(y), y_ref)
    check("dw", conv.weight.grad, wr.grad)
    check("db", conv.bias.grad, br.grad)
    check("dimg", cimgs[0].grad, xin.grad[:, :3])


@pytest.mark.parametrize("cin,cout,k,p,h,w", [(64, 3, 1, 0, 16, 16), (64, 3, 3, 1, 8, 12), (512, 1, 4, 1, 8, 8)])
def test_small_out_conv_sigmoid(cin, cout, k, p, h, w):
    from vae_gan_mark_b200 import layers as L
    torch.manual_seed(7)
    n = 3
    x = bf(torch.randn(n, cin, h, w))
    conv = nn.Conv2d(cin, cout, k, 1, p)
    rx = x.clone().requires_grad_(True)
    y_ref = torch.sigmoid(conv(rx))
    g = torch.randn_like(y_ref)
    y_ref.backward(g)
    convc = nn.Conv2d(cin, cout, k, 1, p).cuda()
    convc.load_state_dict(conv.state_dict())
    xc = nhwc(x)
    y = L.SigmoidOutFn.apply(L.SmallOutConvFn.apply(xc, convc.weight, convc.bias, p))
    y.backward(g.cuda())
    check("y", y, y_ref, 1e-4)
    check("dx", nchw(xc.grad), rx.grad)
    check("dw", convc.weight.grad, conv.weight.grad, 1e-4)
    check("db", convc.bias.grad, conv.bias.grad, 1e-4)


def test_losses():
    from vae_gan_mark_b200 import layers as L
    torch.manual_seed(8)
    a, b = torch.rand(4, 3, 16, 16), torch.rand(4, 3, 16, 16)
    ra = a.clone().requires_grad_(True)
    l_ref = F.l1_loss(ra, b)
    (2.5 * l_ref).backward()
    ac = a.cuda().requires_grad_(True)
    l = L.l1_loss(ac, b.cuda())
    (2.5 * l).backward()
    check("l1", l, l_ref, 1e-5)
    check("l1 grad", ac.grad, ra.grad, 1e-6)
    p = torch.randn(4, 1, 7, 7) * 1.5
    for target in (1, 0, None):
        rp = p.clone().requires_grad_(True)
        ref = (F.relu(1.0 - rp).mean() if target == 1 else F.relu(1.0 + rp).mean() if target == 0 else -rp.mean())
        (0.5 * ref).backward()
        pc = p.cuda().requires_grad_(True)
        got = L.hinge_loss(pc, target)
        (0.5 * got).backward()
        check(f"hinge {target}", got, ref, 1e-5)
        check(f"hinge {target} grad", pc.grad, rp.grad, 1e-6)


@pytest.mark.parametrize("training", [True, False])
def test_spectral_norm_conv(training):
    """SN-Conv4x4 s2: sigma, u, v updates and the gradient through W/sigma vs torch.nn.utils.spectral_norm."""
    from torch.nn.utils import spectral_norm
    from vae_gan_mark_b200 import modules as M
    torch.manual_seed(9)
    cin, cout, n, h, w = 64, 128, 2, 16, 16
    ref = spectral_norm(nn.Conv2d(cin, cout, 4, 2, 1))
    mine = spectral_norm(nn.Conv2d(cin, cout, 4, 2, 1))
    mine.load_state_dict(ref.state_dict())
    ref.weight_orig.data = bf(ref.weight_orig.data)
    mine.weight_orig.data = bf(mine.weight_orig.data)
    ref.train(training); mine.train(training)
    mine = mine.cuda()
    x = bf(torch.randn(n, cin, h, w))
    rx = x.clone().requires_grad_(True)
    y_ref = ref(rx)
    g = bf(torch.randn_like(y_ref))
    y_ref.backward(g)
    xc = nhwc(x)
    sn = M._SNCall(mine, training)
    y = M.run_conv(mine, xc, sn=sn, weight=mine.weight_orig)
    y.backward(nhwc(g).detach())
    check("y", nchw(y), y_ref)
    check("u", mine.weight_u, ref.weight_u, 1e-4)
    check("v", mine.weight_v, ref.weight_v, 1e-4)
    check("dx", nchw(xc.grad), rx.grad)
    check("dw_orig", mine.weight_orig.grad, ref.weight_orig.grad)
    check("db", mine.bias.grad, ref.bias.grad)


def test_fused_adam_and_clip_match_torch():
    from vae_gan_mark_b200.train import FusedAdam
    torch.manual_seed(10)
    shapes = [(64, 32, 3, 3), (128,), (7, 5), (1,)]
    ps_ref = [torch.randn(s).requires_grad_(True) for s in shapes]
    ps = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ps_ref]
    opt_ref = torch.optim.Adam(ps_ref, lr=1e-4, betas=(0.5, 0.999))
    opt = FusedAdam(ps, lr=1e-4, betas=(0.5, 0.999))
    for it in range(3):
        for p, q in zip(ps_ref, ps):
            g = torch.randn_like(p) * (3.0 if it == 0 else 0.01)
            p.grad = g.clone()
            q.grad = g.clone().cuda()
        n_ref = torch.nn.utils.clip_grad_norm_(ps_ref, 1.0)
        opt_ref.step()
        opt.step(max_norm=1.0)
        check(f"norm {it}", opt.norm_sq.sqrt(), n_ref, 1e-5)
        for i, (p, q) in enumerate(zip(ps_ref, ps)):
            check(f"param {it}.{i}", q, p, 1e-6)
            check(f"clipped grad {it}.{i}", q.grad, p.grad, 1e-5)
    sd = opt.state_dict()
    ref_sd = opt_ref.state_dict()
    assert set(sd["state"].keys()) == set(ref_sd["state"].keys())
    check("exp_avg", sd["state"][0]["exp_avg"], ref_sd["state"][0]["exp_avg"], 1e-5)


def test_multi_tensor_adam_over_a_long_ragged_parameter_list():
    """The multi-tensor kernels start tensor t at block (t * 37) mod grid (vg_loss_optim.cu: rotated_tid), so every element
    of every tensor must still be visited exactly once: 60 tensors from 1 element to 3 M elements (several sweeps of the
    grid), odd sizes (scalar tails), against torch's clip_grad_norm_ + Adam for two steps."""
    from vae_gan_mark_b200.train import FusedAdam
    g = torch.Generator().manual_seed(77)
    sizes = [1, 3, 4, 5, 255, 256, 257, 1023, 4099, 65537, 151552 * 4, 151552 * 4 + 4, 606211, 1000003, 3000001]
    sizes = sizes + [int(torch.randint(1, 200000, (1,), generator=g)) for _ in range(45)]
    ps_ref = [torch.randn(n, generator=g).requires_grad_(True) for n in sizes]
    ps = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ps_ref]
    opt_ref = torch.optim.Adam(ps_ref, lr=1e-3, betas=(0.5, 0.999))
    opt = FusedAdam(ps, lr=1e-3, betas=(0.5, 0.999))
    for it in range(2):
        for p, q in zip(ps_ref, ps):
            gr = torch.randn(p.shape, generator=g) * (0.02 if it == 0 else 1e-4)
            p.grad = gr.clone()
            q.grad = gr.clone().cuda()
        n64 = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in ps_ref))     # torch's own fp32 norm is 2e-5 off at 11 M elements
        torch.nn.utils.clip_grad_norm_(ps_ref, 1.0)
        opt_ref.step()
        opt.step(max_norm=1.0)
        check(f"norm {it}", opt.norm_sq.sqrt(), n64, 3e-5)
        for i, (p, q) in enumerate(zip(ps_ref, ps)):
            assert float((q.detach().cpu() - p.detach()).abs().max()) <= 2e-6, (it, i, sizes[i])
            check(f"clipped grad {it}.{i}", q.grad, p.grad, 1e-4)      # the two clip coefficients differ by the norms' 2e-5
            # (1 - beta2) is formed in fp32 by the kernel and in double by torch: 1.3e-5 relative on exp_avg_sq
            check(f"exp_avg_sq {it}.{i}", opt.exp_avg_sq[i], opt_ref.state[p]["exp_avg_sq"], 2e-4)


@pytest.mark.parametrize("n,offset", [(1000003, 0), (4096, 0), (1000003, 1), (7, 0)])
def test_single_tensor_adam_step_matches_torch(n, offset):
    """vg_adam_step / vg_sumsq, the flat-buffer entry points of the C ABI: 16-byte accesses when the four arrays are
    aligned (offset 0), the scalar loop otherwise (views starting one element into their buffers) and for the tail."""
    from vae_gan_mark_b200 import ops
    from vae_gan_mark_b200._lib import VgError
    g = torch.Generator().manual_seed(n + offset)
    p_ref = torch.randn(n, generator=g).requires_grad_(True)
    opt_ref = torch.optim.Adam([p_ref], lr=1e-3, betas=(0.5, 0.999))
    bufs = [torch.zeros(n + offset, device="cuda") for _ in range(4)]
    p, gr, m, v = (b[offset:] for b in bufs)
    p.copy_(p_ref.detach())
    nsq = torch.zeros((), device="cuda")
    for it in range(1, 3):
        grad = torch.randn(n, generator=g) * (0.05 if it == 1 else 1e-4)
        p_ref.grad = grad.clone()
        gr.copy_(grad)
        n64 = (grad.double() ** 2).sum().sqrt()
        torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt_ref.step()
        if offset:      # vg_sumsq reads 16 bytes at a time and refuses other buffers with an error code
            with pytest.raises(VgError):
                ops.sumsq(gr, nsq)
            nsq.copy_((gr.double() ** 2).sum())
        else:
            ops.sumsq(gr, nsq)
        ops.adam_step(p, gr, m, v, 1e-3, 0.5, 0.999, 1e-8, it, nsq, 1.0, True)
        check(f"norm {it}", nsq.sqrt(), n64, 3e-5)
        assert float((p.cpu() - p_ref.detach()).abs().max()) <= 2e-6
        check(f"clipped grad {it}", gr, p_ref.grad, 1e-4)
        check(f"exp_avg {it}", m, opt_ref.state[p_ref]["exp_avg"], 1e-4)
        # (1 - beta2) is formed in fp32 by the kernel and in double by torch: 1.3e-5 relative on exp_avg_sq
        check(f"exp_avg_sq {it}", v, opt_ref.state[p_ref]["exp_avg_sq"], 2e-4)


def test_strided_copy_layouts_exact():
    """vg_strided_copy (tiled through shared memory) against torch for the re-layouts the step uses: weight OIHW ->
    GEMM layouts, gradients back, NCHW <-> NHWC, channel-slice destinations, broadcast sources, scale, accumulate.
    fp32 -> fp32 must be bit-exact; bf16 destinations must equal torch's round-to-nearest conversion."""
    from vae_gan_mark_b200 import ops
    g = torch.Generator().manual_seed(5)
    cases = [((512, 256, 3, 3), (0, 2, 3, 1)), ((512, 256, 3, 3), (2, 3, 1, 0)), ((128, 3, 3, 200), (0, 3, 1, 2)),
             ((64, 128, 4, 4), (1, 2, 3, 0)), ((3, 5, 37, 41), (0, 2, 3, 1)), ((3, 37, 41, 8), (0, 3, 1, 2)),
             ((100003,), (0,)), ((2, 3, 4, 5, 6), (4, 2, 0, 3, 1)), ((1024, 1024, 2, 2), (0, 2, 3, 1))]
    for shape, perm in cases:
        src = torch.randn(shape, generator=g).cuda()
        view = src.permute(perm)
        for dt in (torch.float32, torch.bfloat16):
            out = torch.full(view.shape, float("nan"), dtype=dt, device="cuda")
            ops.strided_copy(view, out)
            assert torch.equal(out, view.to(dt)), (shape, perm, dt)
        srcb = src.to(torch.bfloat16).permute(perm)
        out = torch.empty(view.shape, dtype=torch.float32, device="cuda")
        ops.strided_copy(srcb, out)
        assert torch.equal(out, srcb.float()), (shape, perm, "bf16->fp32")
    # channel-slice destination + broadcast source
    z = torch.randn(4, 24, generator=g).cuda()
    buf = torch.zeros(4, 1, 6, 40, device="cuda", dtype=torch.bfloat16)
    ops.strided_copy(z.view(4, 1, 1, 24).expand(4, 1, 6, 24), buf[..., 8:32])
    assert torch.equal(buf[..., 8:32], z.to(torch.bfloat16).view(4, 1, 1, 24).expand(4, 1, 6, 24))
    assert not buf[..., :8].any() and not buf[..., 32:].any()
    # device scalar (inverse) scale as used for W / sigma
    w = torch.randn(96, 40, 4, 4, generator=g).cuda()
    sigma = torch.tensor([1.7], device="cuda")
    out = torch.empty(96, 4, 4, 40, device="cuda")
    ops.strided_copy(w.permute(0, 2, 3, 1), out, sigma, scale_inverse=True)
    torch.testing.assert_close(out, w.permute(0, 2, 3, 1) / 1.7, rtol=1e-6, atol=0)


@pytest.mark.parametrize("batch,steps,precision", [(5, 7, "fp32"), (16, 60, "fp32"), (64, 3, "fp32"), (16, 60, "bf16")])
def test_gru_text_encoder_matches_torch_gru(batch, steps, precision):
    """CharacterTokenEncoder's cluster-kernel biGRU (vg_gru.cu: one launch per layer walks all time steps) against
    torch.nn.GRU in float64 on the CPU: outputs, input gradient and every parameter gradient.  The recurrence is
    all-fp32 FMA; the time-parallel GEMMs around it run on the tcgen05 kernels with split-bf16 operands in BOTH precision
    modes (layers.GRULayerFn), the embedding and the pooling on vg_text.cu: tolerance 1e-4 relative L2 (the reference's
    cuDNN GRU uses TF32 GEMMs, ~1e-3)."""
    import vae_gan_mark_b200
    from vae_gan_mark_b200 import modules as M
    vae_gan_mark_b200.set_precision(precision)
    try:
        _gru_case(M, batch, steps, 1e-4)
    finally:
        vae_gan_mark_b200.set_precision("bf16")


def _gru_case(M, batch, steps, tol):
    torch.manual_seed(batch * 100 + steps)
    enc = M.CharacterTokenEncoder(M.ALPHABET_STR, 128, 256, 2, 8).cuda().train()
    enc.rnn.dropout = 0.0
    ref = nn.GRU(128, 256, num_layers=2, batch_first=True, bidirectional=True).double()
    ref.load_state_dict({k: v.detach().cpu().double() for k, v in enc.rnn.state_dict().items()})
    idx = torch.randint(0, enc.vocab_size, (batch, steps))
    emb_w = enc.embedding.weight.detach().cpu().double().requires_grad_(True)
    out_ref, _ = ref(F.embedding(idx, emb_w, padding_idx=0))
    y_ref = F.adaptive_avg_pool1d(out_ref.permute(0, 2, 1), 8).unsqueeze(2)
    gy = torch.randn(y_ref.shape, dtype=torch.float64)
    y_ref.backward(gy)
    y = enc(idx.cuda())
    y.backward(gy.float().cuda())
    check("gru out", y, y_ref, tol)
    check("gru d embedding", enc.embedding.weight.grad, emb_w.grad, tol)
    for (name, p), (_, pr) in zip(enc.rnn.named_parameters(), ref.named_parameters()):
        check(f"gru d {name}", p.grad, pr.grad, tol)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 24, 64, 64), (3, 40, 20, 128, 64), (2, 16, 8, 64, 128), (1, 50, 30, 64, 32),
                                            (2, 17, 9, 192, 64)])
def test_halo_mode_conv_equals_per_tap_path(n, h, w, cin, cout):
    """fprop halo mode (the 18 x 10 activation halo of a 16 x 8 pixel tile is loaded once per 64-channel chunk and the
    nine taps are nine shifted UMMA descriptors into it; weights resident in shared memory when they fit) must
    reproduce the per-tap path: same products, same fp32 accumulator -- only the order of the K loop differs."""
    from vae_gan_mark_b200 import conv
    from vae_gan_mark_b200.conv import ConvLinear
    torch.manual_seed(3)
    op = ConvLinear(cin, cout, 3, 3, 1, (1, 1))
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    bias = torch.randn(cout, device="cuda")
    wf = op.prep_fwd(wt)
    outs = {}
    for mode in (-1, 1):
        conv.HALO_MODE = mode
        try:
            outs[mode] = op.forward(x, wf, bias, 1).float()
        finally:
            conv.HALO_MODE = 0
    check("halo vs per-tap", outs[1], outs[-1], 1e-4)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), bf(wt.cpu()).cuda(), bias, padding=1)).permute(0, 2, 3, 1)
    check("halo vs torch", outs[1], ref, TOL)


def test_text_front_end_kernels():
    """Device tokeniser (code point -> index lookup table) against the reference's host loop (vae-gan-v2.py:89-100), the
    embedding gather / per-row gradient sum against F.embedding, and the sequence pooling into the NHWC text map against
    nn.AdaptiveAvgPool1d (vae-gan-v2.py:107-113), forward and backward; also through the whole module with strings."""
    from vae_gan_mark_b200 import layers as L, modules as M, ops
    torch.manual_seed(11)
    enc = M.CharacterTokenEncoder(M.ALPHABET_STR_UNET, 128, 256, 2, 28).cuda().train()
    enc.rnn.dropout = 0.0
    texts = ["SALE", "Привет, мир! ёЁ", "", "x" * 100, "日本語 abc ~", "tab\there", "Free shipping on orders over $25"]
    idx = enc.indices(texts, 60)
    assert idx.dtype == torch.long and idx.is_cuda
    assert torch.equal(idx.cpu(), enc.tokens_to_indices(texts, 60))
    # embedding
    w = enc.embedding.weight.detach().clone().requires_grad_()
    w_ref = w.detach().clone().requires_grad_()
    e = L.EmbeddingFn.apply(idx, w, 0)
    e_ref = F.embedding(idx, w_ref, padding_idx=0)
    g = torch.randn_like(e_ref)
    e.backward(g)
    e_ref.backward(g)
    assert torch.equal(e, e_ref)
    check("embedding dW", w.grad, w_ref.grad, 1e-6)
    assert float(w.grad[0].abs().max()) == 0.0
    # pooling: fp32 sequence -> NHWC map (fp32 exactly, bf16 to rounding), every bin layout incl. overlapping bins
    for l, wout in ((60, 28), (60, 8), (60, 4), (7, 16), (60, 60), (60, 1)):
        seq = torch.randn(3, l, 512, device="cuda", requires_grad=True)
        seq_ref = seq.detach().clone().requires_grad_()
        y = L.SeqPoolFn.apply(seq, wout, torch.float32)
        y_ref = F.adaptive_avg_pool1d(seq_ref.permute(0, 2, 1), wout).permute(0, 2, 1).unsqueeze(1)
        gy = torch.randn_like(y_ref)
        y.backward(gy)
        y_ref.backward(gy)
        check(f"seqpool {l}->{wout}", y, y_ref, 1e-6)
        check(f"seqpool {l}->{wout} dseq", seq.grad, seq_ref.grad, 1e-6)
        yb = L.SeqPoolFn.apply(seq.detach(), wout, torch.bfloat16)
        check(f"seqpool {l}->{wout} bf16", yb, y_ref, 4e-3)
    # whole module on strings == on the host-tokenised indices; gradient reaches the embedding
    y1 = enc(texts)
    y2 = enc(enc.tokens_to_indices(texts, 60).cuda())
    assert tuple(y1.shape) == (len(texts), 512, 1, 28)
    check("strings vs host-tokenised indices", y1, y2, 1e-6)       # (not bit-equal: split-K partial tiles meet in atomics)
    y1.sum().backward()
    assert enc.embedding.weight.grad is not None and float(enc.embedding.weight.grad.abs().max()) > 0
    t = enc.nhwc_features(texts)
    assert tuple(t.shape) == (len(texts), 1, 28, 512) and t.dtype == ops.act_dtype()
    check("nhwc text map", t.float().permute(0, 3, 1, 2), y1, 4e-3)
