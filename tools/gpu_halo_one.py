import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vae_gan_mark_b200 import conv
from vae_gan_mark_b200.conv import ConvLinear
n, h, w, cin, cout = 64, 128, 128, 64, 64
op = ConvLinear(cin, cout, 3, 3, 1, (1, 1))
x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
wf = op.prep_fwd(torch.randn(cout, cin, 3, 3, device="cuda") * 0.05)
out = conv.new_act(n, h, w, cout, "cuda")
conv.HALO_MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(3):
    op.forward(x, wf, out=out)
torch.cuda.synchronize()
print("ok")
