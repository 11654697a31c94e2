"""Time the text encoder (B=64, T=60, 2-layer biGRU 256) forward+backward: cluster-kernel path vs stock cuDNN (dev tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vae_gan_mark_b200 import modules as M, ops  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def kernels_only():
    from vae_gan_mark_b200 import _lib
    L = _lib.lib()
    print("max active clusters fwd bg8/bg16, bwd bg8/bg16:", L.vg_gru_max_active_clusters(0, 8),
          L.vg_gru_max_active_clusters(0, 16), L.vg_gru_max_active_clusters(1, 8), L.vg_gru_max_active_clusters(1, 16))
    b, t, h = 64, 60, 256
    xproj = torch.randn(b, t, 2, 3 * h, device="cuda")
    w_hh = torch.randn(2, 3 * h, h, device="cuda") * 0.05
    b_hh = torch.randn(2, 3 * h, device="cuda") * 0.05
    out = torch.empty(b, t, 2 * h, device="cuda")
    gates = torch.empty(2, b, t, 4, h, device="cuda")
    dgx = torch.empty(b, t, 2, 3 * h, device="cuda")
    dgh = torch.empty(2, b, t, 3 * h, device="cuda")
    dout = torch.randn(b, t, 2 * h, device="cuda")
    print("gru_seq_fwd kernel ms", timeit(lambda: ops.gru_seq_fwd(xproj, w_hh, b_hh, out, gates), 5, 2))
    print("gru_seq_bwd kernel ms", timeit(lambda: ops.gru_seq_bwd(dout, out, gates, w_hh, dgx, dgh), 5, 2))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "kernels":
        return kernels_only()
    torch.manual_seed(0)
    enc = M.CharacterTokenEncoder(M.ALPHABET_STR, 128, 256, 2, 8).cuda().train()
    idx = torch.randint(0, enc.vocab_size, (64, 60), device="cuda")
    gy = torch.randn(64, 512, 1, 8, device="cuda")

    def ours():
        y = enc(idx)
        y.backward(gy)

    def stock():
        out, _ = enc.rnn(enc.embedding(idx))
        y = enc.adaptive_pool(out.permute(0, 2, 1)).unsqueeze(2)
        y.backward(gy)
    print("ours  eager fwd+bwd ms", timeit(ours))
    print("stock eager fwd+bwd ms", timeit(stock))
    for name, fn in (("ours", ours), ("stock", stock)):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
        torch.cuda.synchronize()
        print(name, "graph fwd+bwd ms", timeit(g.replay))
    # kernels alone
    b, t, h = 64, 60, 256
    xproj = torch.randn(b, t, 2, 3 * h, device="cuda")
    w_hh = torch.randn(2, 3 * h, h, device="cuda") * 0.05
    b_hh = torch.randn(2, 3 * h, device="cuda") * 0.05
    out = torch.empty(b, t, 2 * h, device="cuda")
    gates = torch.empty(2, b, t, 4, h, device="cuda")
    dgx = torch.empty(b, t, 2, 3 * h, device="cuda")
    dgh = torch.empty(2, b, t, 3 * h, device="cuda")
    dout = torch.randn(b, t, 2 * h, device="cuda")
    print("gru_seq_fwd kernel ms", timeit(lambda: ops.gru_seq_fwd(xproj, w_hh, b_hh, out, gates)))
    print("gru_seq_bwd kernel ms", timeit(lambda: ops.gru_seq_bwd(dout, out, gates, w_hh, dgx, dgh)))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ours()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
