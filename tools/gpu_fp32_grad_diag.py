"""Per-tensor gradient error of one fp32-mode training step against the float64 oracle, listed in backward order
(dev tool: shows at which layer of the backward pass a precision loss enters).
usage: python tools/gpu_fp32_grad_diag.py <family> <h> <w> <batch> <z> [dirty]
``dirty``: first fill 6 GB of the caching allocator's pool with large finite garbage, so that a kernel that reads memory
it never wrote (and gets zeros from a fresh cudaMalloc in a clean process) shows up."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle.step import LossWeights as OLW, make_optimizers, synthetic_batch, train_step  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def main():
    import vae_gan_mark_b200 as vg
    from test_step_parity_gpu import build_pair
    from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer
    family = sys.argv[1] if len(sys.argv) > 1 else "oldv"
    h, w, batch, z = (int(v) for v in (sys.argv[2:6] if len(sys.argv) > 5 else (32, 64, 2, 128)))
    vg.set_precision("fp32")
    if "dirty" in sys.argv:
        junk = [torch.full((256, 1024, 1024), 1000.0, device="cuda") for _ in range(6)]
        junk += [torch.full((1 << (10 + k),), 1000.0, device="cuda") for k in range(16) for _ in range(8)]
        torch.cuda.synchronize()
        del junk
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    og, od, mg, md = build_pair(family, h, w, z)
    og64, od64 = copy.deepcopy(og).double(), copy.deepcopy(od).double()
    wts = OLW.for_family(family)
    ru, en, mask, texts = synthetic_batch(batch, h, w, step=0)
    eps = torch.randn(batch, z, 1, 1, generator=torch.Generator().manual_seed(5))
    orig = torch.randn_like
    try:
        torch.randn_like = lambda t, **k: eps.to(t.dtype) if tuple(t.shape) == tuple(eps.shape) else orig(t, **k)
        ref32 = train_step(og, od, *make_optimizers(og, od), (ru, en, mask, texts), wts)
        ref = train_step(og64, od64, *make_optimizers(og64, od64), (ru.double(), en.double(), mask.double(), texts), wts)
    finally:
        torch.randn_like = orig
    grads = {}
    trainer = VAEGANTrainer(mg, md, LossWeights(wts.recon, wts.kl, wts.gan),
                            grad_hook=lambda which, params: grads.setdefault(which, [p.grad.clone() if p.grad is not None else None for p in params]))
    enc = getattr(mg, "style_vae_encoder_module", None) or mg.encoder
    enc.__dict__["eps_fn"] = lambda shape: eps.clone()
    out = trainer.step(ru.cuda(), en.cuda(), mask.cuda(), texts)
    torch.cuda.synchronize()
    print("fake", rel(out["fake"], ref.recon), "mu", rel(out["mu"], ref.mu))
    for which, mod, rg, rg32 in (("D", md, ref.d_grads, ref32.d_grads), ("G", mg, ref.g_grads, ref32.g_grads)):
        rows = [(n, g) for (n, _), g in zip(mod.named_parameters(), grads[which]) if g is not None and n in rg]
        for n, g in reversed(rows):
            print(f"{which}.{n:70s} ours {rel(g, rg[n]):.1e}  oracle32 {rel(rg32[n], rg[n]):.1e}  |g| {float(rg[n].norm()):.1e}")


if __name__ == "__main__":
    main()
