import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
from vae_gan_mark_b200 import modules as M, layers as L, ops, _lib
from cuda.bindings import driver as cu

def ctx():
    err, c = cu.cuCtxGetCurrent()
    return err, int(c) if c is not None else None

x = torch.randn(3, 16, 16, 64).to(torch.bfloat16).cuda()
print("after torch init", ctx())
w = torch.randn(64, 64, 3, 3).cuda()
out = torch.empty(64, 3, 3, 64, dtype=torch.bfloat16, device="cuda")
ops.strided_copy(w.permute(0, 2, 3, 1), out)
torch.cuda.synchronize()
print("after strided_copy", ctx())
conv = nn.Conv2d(64, 64, 3, padding=1).cuda()
try:
    y = M.run_conv(conv, x)
    torch.cuda.synchronize()
    print("conv ok", y.shape)
except Exception as e:
    print("conv failed:", e)
print("after conv", ctx())
import ctypes as C
sm = C.c_int(); a = C.c_int(); b = C.c_int()
print("device_info rc", _lib.lib().vg_device_info(C.byref(sm), C.byref(a), C.byref(b)), sm.value, a.value, b.value)
try:
    y = M.run_conv(conv, x)
    torch.cuda.synchronize()
    print("conv (2nd try) ok", y.shape)
except Exception as e:
    print("conv 2nd failed:", e)
