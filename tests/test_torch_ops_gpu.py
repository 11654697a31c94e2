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
