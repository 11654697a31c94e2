"""The torch.library registration (vae_gan_mark_b200/torch_ops.py): torch.ops.vaegan.* custom ops with fake
implementations and registered autograd, checked against plain PyTorch fp32 on bf16-rounded inputs (relative L2 <= 1e-2,
as in test_ops_gpu.py) and with torch.library.opcheck (schema + fake-tensor consistency)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def bf(t):
    return t.to(torch.bfloat16).float()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda().requires_grad_(True)


def nchw(t):
    return t.detach().float().cpu().permute(0, 3, 1, 2)


@pytest.mark.parametrize("cin,cout,k,s,p,h,w,act", [(64, 128, 3, 1, 1, 16, 12, 1), (128, 64, 4, 2, 1, 16, 16, 2), (256, 64, 1, 1, 0, 8, 8, 0)])
def test_conv2d_op(cin, cout, k, s, p, h, w, act):
    import vae_gan_mark_b200.torch_ops  # noqa: F401  (registers the ops)
    torch.manual_seed(0)
    x = bf(torch.randn(2, cin, h, w))
    wt = bf(torch.randn(cout, cin, k, k) * 0.05)
    b = torch.randn(cout)
    rx, rw, rb = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = F.conv2d(rx, rw, rb, s, p)
    y_ref = F.relu(y_ref) if act == 1 else (F.leaky_relu(y_ref, 0.2) if act == 2 else y_ref)
    gy = bf(torch.randn_like(y_ref))
    y_ref.backward(gy)
    xc, wc, bc = nhwc(x), wt.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = torch.ops.vaegan.conv2d(xc, wc, bc, s, p, p, act)
    y.backward(nhwc(gy).detach())
    assert rel(nchw(y), y_ref) <= TOL
    assert rel(nchw(xc.grad), rx.grad) <= TOL
    assert rel(wc.grad, rw.grad) <= TOL
    assert rel(bc.grad, rb.grad) <= TOL
    torch.library.opcheck(torch.ops.vaegan.conv2d.default, (xc.detach(), wc.detach(), bc.detach(), s, p, p, act),
                          test_utils=("test_schema", "test_faketensor"))


def test_conv_transpose2d_batch_norm_film_ops():
    import vae_gan_mark_b200.torch_ops  # noqa: F401
    torch.manual_seed(1)
    n, cin, cout, h, w = 2, 128, 64, 8, 8
    x = bf(torch.randn(n, cin, h, w))
    wt = bf(torch.randn(cin, cout, 2, 2) * 0.05)
    gamma, beta = torch.rand(cout) + 0.5, torch.randn(cout) * 0.2
    gb = bf(torch.randn(n, 2 * cout, 2 * h, 2 * w))
    leaves = [t.clone().requires_grad_(True) for t in (x, wt, gamma, beta, gb)]
    rx, rw, rg, rbeta, rgb = leaves
    u = F.conv_transpose2d(rx, rw, None, 2, 0)
    v = F.relu(F.batch_norm(u, None, None, rg, rbeta, True, 0.1, 1e-5))
    y_ref = rgb[:, :cout] * v + rgb[:, cout:]
    gy = bf(torch.randn_like(y_ref))
    y_ref.backward(gy)
    xc, wc = nhwc(x), wt.cuda().requires_grad_(True)
    gc, bc, gbc = gamma.cuda().requires_grad_(True), beta.cuda().requires_grad_(True), nhwc(gb)
    u2 = torch.ops.vaegan.conv_transpose2d(xc, wc, None, 2, 0, 2 * h, 2 * w, 0)
    v2, _ = torch.ops.vaegan.batch_norm_act(u2, gc, bc, 1e-5, 1)
    y = torch.ops.vaegan.film(gbc, v2)
    y.backward(nhwc(gy).detach())
    assert rel(nchw(y), y_ref) <= TOL
    assert rel(nchw(xc.grad), rx.grad) <= 2 * TOL
    assert rel(wc.grad, rw.grad) <= 2 * TOL
    assert rel(gc.grad, rg.grad) <= 2 * TOL and rel(bc.grad, rbeta.grad) <= 2 * TOL
    assert rel(nchw(gbc.grad), rgb.grad) <= TOL


def test_resize_gate_reparam_and_loss_ops():
    """The remaining stateless primitives of the step as dispatcher ops: bilinear resize of an NHWC map, per-channel
    gate (vae-gan-oldv.py:165-176, 226-231), reparameterisation + KL and the L1 / hinge losses (vae-gan.py:133-136,
    313-320, 419-420) -- forward and registered autograd against plain PyTorch, plus opcheck."""
    import vae_gan_mark_b200.torch_ops  # noqa: F401
    torch.manual_seed(2)
    n, c = 2, 64
    # bilinear resize
    t = bf(torch.randn(n, c, 4, 4))
    rt = t.clone().requires_grad_(True)
    u_ref = F.interpolate(rt, size=(8, 16), mode="bilinear", align_corners=False)
    gu = bf(torch.randn_like(u_ref))
    u_ref.backward(gu)
    tc = nhwc(t)
    u = torch.ops.vaegan.upsample_bilinear2d(tc, 8, 16)
    u.backward(nhwc(gu).detach())
    assert rel(nchw(u), u_ref) <= TOL and rel(nchw(tc.grad), rt.grad) <= TOL
    torch.library.opcheck(torch.ops.vaegan.upsample_bilinear2d.default, (tc.detach(), 8, 16),
                          test_utils=("test_schema", "test_faketensor"))
    # per-channel gate
    x, s = bf(torch.randn(n, c, 8, 8)), torch.rand(c) + 0.2
    rx, rs = x.clone().requires_grad_(True), s.clone().requires_grad_(True)
    y_ref = rx * rs.view(1, c, 1, 1)
    gy = bf(torch.randn_like(y_ref))
    y_ref.backward(gy)
    xc, sc = nhwc(x), s.cuda().requires_grad_(True)
    y = torch.ops.vaegan.channel_gate(xc, sc)
    y.backward(nhwc(gy).detach())
    assert rel(nchw(y), y_ref) <= TOL and rel(nchw(xc.grad), rx.grad) <= TOL and rel(sc.grad, rs.grad) <= 1e-4
    torch.library.opcheck(torch.ops.vaegan.channel_gate.default, (xc.detach(), sc.detach()),
                          test_utils=("test_schema", "test_faketensor"))
    # reparameterisation + KL (fp32)
    b, z = 5, 128
    heads, bm, bl, eps = torch.randn(b, 2 * z) * 0.5, torch.randn(z) * 0.1, torch.randn(z) * 0.1, torch.randn(b, z)
    rh, rbm, rbl = (v.clone().requires_grad_(True) for v in (heads, bm, bl))
    mu_ref, lv_ref = rh[:, :z] + rbm, rh[:, z:] + rbl
    z_ref = mu_ref + eps * torch.exp(0.5 * lv_ref)
    kl_ref = torch.mean(-0.5 * torch.mean(1 + lv_ref - mu_ref.pow(2) - lv_ref.exp(), dim=1))
    gz = torch.randn(b, z)
    (0.37 * kl_ref + (z_ref * gz).sum() + 0.1 * mu_ref.sum()).backward()
    hc, bmc, blc = (v.cuda().requires_grad_(True) for v in (heads, bm, bl))
    mu, lv, zz, kl = torch.ops.vaegan.reparam_kl(hc, bmc, blc, eps.cuda())
    (0.37 * kl + (zz * gz.cuda()).sum() + 0.1 * mu.sum()).backward()
    for got, want in ((mu, mu_ref), (lv, lv_ref), (zz, z_ref), (kl, kl_ref), (hc.grad, rh.grad), (bmc.grad, rbm.grad),
                      (blc.grad, rbl.grad)):
        assert rel(got, want) <= 1e-5
    torch.library.opcheck(torch.ops.vaegan.reparam_kl.default, (hc.detach(), bmc.detach(), blc.detach(), eps.cuda()),
                          test_utils=("test_schema", "test_faketensor"))
    # losses (fp32)
    a, bb = torch.rand(4, 3, 16, 16), torch.rand(4, 3, 16, 16)
    ra = a.clone().requires_grad_(True)
    (2.5 * F.l1_loss(ra, bb)).backward()
    ac = a.cuda().requires_grad_(True)
    l1 = torch.ops.vaegan.l1_loss(ac, bb.cuda())
    (2.5 * l1).backward()
    assert rel(l1, F.l1_loss(a, bb)) <= 1e-5 and rel(ac.grad, ra.grad) <= 1e-5
    p = torch.randn(4, 1, 7, 7) * 1.5
    for mode, fn in ((1, lambda q: F.relu(1.0 - q).mean()), (0, lambda q: F.relu(1.0 + q).mean()), (2, lambda q: -q.mean())):
        rp = p.clone().requires_grad_(True)
        (1.7 * fn(rp)).backward()
        pc = p.cuda().requires_grad_(True)
        hl = torch.ops.vaegan.hinge_loss(pc, mode)
        (1.7 * hl).backward()
        assert rel(hl, fn(p)) <= 1e-5 and rel(pc.grad, rp.grad) <= 1e-5, mode
    torch.library.opcheck(torch.ops.vaegan.hinge_loss.default, (pc.detach(), 1), test_utils=("test_schema", "test_faketensor"))
