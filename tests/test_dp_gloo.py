"""Host-side logic of the data-parallel path on CPU: world_size 2, gloo backend (no GPU needed).

Checks that DataParallelReducer averages gradients bucket by bucket exactly like a single process that saw the
concatenated batch (all loss terms are batch means), that parameters are broadcast from rank 0, and that the
reverse-order bucketing covers every parameter once.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker_overlap(rank, world, port, ret):
    """Same check with the all-reduces issued from autograd hooks during the backward, over two consecutive steps,
    with one parameter that receives no gradient (its bucket must still be reduced by hook())."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vae_gan_mark_b200.parallel import DataParallelReducer
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    unused = torch.nn.Parameter(torch.zeros(3))
    params = list(net.parameters()) + [unused]
    red = DataParallelReducer(world, bucket_bytes=300)
    red.install_hooks(params)
    out = []
    for step in range(2):
        for p in params:
            p.grad = None
        g = torch.Generator().manual_seed(70 + step)
        x_all, y_all = torch.randn(6, 8, generator=g), torch.randn(6, 4, generator=g)
        xs, ys = x_all[rank * 3:(rank + 1) * 3], y_all[rank * 3:(rank + 1) * 3]
        torch.nn.functional.mse_loss(net(xs), ys).backward()
        launched_during_backward = sum(st["launched"] for st in red._states)
        red.hook("G", params)
        ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        ref.load_state_dict(net.state_dict())
        torch.nn.functional.mse_loss(ref(x_all), y_all).backward()
        ok = all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(net.parameters(), ref.parameters()))
        out.append((ok, launched_during_backward, unused.grad is None))
    ret[rank] = out
    dist.destroy_process_group()


def test_two_rank_overlapped_allreduce_from_autograd_hooks():
    world, port = 2, 31000 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_overlap, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        for ok, launched, unused_none in ret[rank]:
            assert ok, "averaged gradients differ from the global-batch gradient"
            assert launched >= 1, "no bucket was launched from the autograd hooks"
            assert unused_none


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vae_gan_mark_b200.parallel import DataParallelReducer
    torch.manual_seed(100 + rank)                      # different init per rank on purpose
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    red = DataParallelReducer(world, bucket_bytes=300)  # tiny buckets -> several all-reduces
    red.broadcast_parameters(list(net.parameters()))
    w0 = [p.detach().clone() for p in net.parameters()]
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(6, 8, generator=g), torch.randn(6, 4, generator=g)
    xs, ys = x_all[rank * 3:(rank + 1) * 3], y_all[rank * 3:(rank + 1) * 3]
    torch.nn.functional.mse_loss(net(xs), ys).backward()
    params = list(net.parameters())
    assert sum(len(b) for b in red.make_buckets(params)) == len(params)
    assert len(red.make_buckets(params)) > 1
    red.hook("G", params)
    ret[rank] = ([p.grad.clone() for p in params], w0)
    dist.destroy_process_group()


def test_two_rank_gradient_average_matches_global_batch():
    world, port = 2, 29000 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (g0, w0), (g1, w1) = ret[0], ret[1]
    for a, b in zip(w0, w1):
        assert torch.equal(a, b), "parameters were not broadcast from rank 0"
    for a, b in zip(g0, g1):
        assert torch.allclose(a, b, atol=1e-7), "ranks disagree after the all-reduce"
    # single-process reference on the full batch with rank 0's weights
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    with torch.no_grad():
        for p, w in zip(net.parameters(), w0):
            p.copy_(w)
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(6, 8, generator=g), torch.randn(6, 4, generator=g)
    torch.nn.functional.mse_loss(net(x_all), y_all).backward()
    for p, a in zip(net.parameters(), g0):
        assert torch.allclose(p.grad, a, atol=1e-6)


def _worker_model_params(rank, world, port, ret):
    """The real parameter sets of the drop-in models (oldv generator incl. the (1,C,1,1) gates, the 4-D positional
    encoding and the Conv1d weight; the discriminator with its spectral-norm weights), conv weights in channels_last
    memory order as VAEGANTrainer keeps them, gradients in the layouts the backward produces."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vae_gan_mark_b200 import modules as M
    from vae_gan_mark_b200.parallel import DataParallelReducer
    from vae_gan_mark_b200.train import weights_channels_last
    torch.manual_seed(3)
    G = M.VAEGAN_UNet_SpatialFiLM_OldV(4, 32, patch_shape=(64, 32))
    D = M.Discriminator(3)
    converted = weights_channels_last(G) + weights_channels_last(D)
    ok, n_cl, n_buckets = True, 0, 0
    for which, params in (("G", list(G.parameters())), ("D", list(D.parameters()))):
        red = DataParallelReducer(world, bucket_bytes=1 << 20)
        n_buckets += len(red.make_buckets(params))
        bases = []
        for k, p in enumerate(params):
            base = torch.randn(p.shape, generator=torch.Generator().manual_seed(1000 + k))
            bases.append(base)
            if k % 3 == 0:
                p.grad = (base * (rank + 1)).contiguous()                    # a fresh contiguous gradient
            else:
                p.grad = torch.empty_like(p).copy_(base * (rank + 1))        # the parameter's own memory order
            n_cl += int(p.dim() == 4 and not p.grad.is_contiguous())
        strides = [p.grad.stride() for p in params]
        red.hook(which, params)
        for p, base, st in zip(params, bases, strides):
            ok = ok and torch.allclose(p.grad, base * 1.5, rtol=1e-6, atol=1e-6) and p.grad.stride() == st
    ret[rank] = (ok, converted, n_cl, n_buckets)
    dist.destroy_process_group()


def test_two_rank_reduction_over_the_real_parameter_sets():
    world, port = 2, 33000 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_model_params, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        ok, converted, n_cl, n_buckets = ret[rank]
        assert ok, "averaged gradients (or their memory layout) are wrong for some parameter"
        assert converted > 20 and n_cl > 10 and n_buckets > 4


def _worker_bf16(rank, world, port, ret):
    """bf16 gradient buckets (the option for the 568 M-parameter configuration): averaged in bf16 on the wire, cast back
    into the fp32 gradients, whose shape and memory order are kept; every rank ends up with the SAME values."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vae_gan_mark_b200.parallel import DataParallelReducer
    torch.manual_seed(9)
    params = [torch.nn.Parameter(torch.zeros(64, 32, 3, 3).contiguous(memory_format=torch.channels_last)),
              torch.nn.Parameter(torch.zeros(64)), torch.nn.Parameter(torch.zeros(10, 7))]
    bases = [torch.randn(p.shape, generator=torch.Generator().manual_seed(50 + k)) for k, p in enumerate(params)]
    for p, b in zip(params, bases):
        p.grad = torch.empty_like(p).copy_(b * (rank + 1))
    red = DataParallelReducer(world, bucket_bytes=4096, grad_dtype=torch.bfloat16)
    red.hook("G", params)
    ok = all(p.grad.dtype == torch.float32 and p.grad.stride() == p.stride() and
             torch.allclose(p.grad, b * 1.5, rtol=2e-2, atol=1e-3) for p, b in zip(params, bases))
    ret[rank] = (ok, [p.grad.clone() for p in params])
    dist.destroy_process_group()


def test_two_rank_bf16_buckets():
    world, port = 2, 35000 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_bf16, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0], "bf16-bucket average is off or changed the gradient layout"
    for a, b in zip(ret[0][1], ret[1][1]):
        assert torch.equal(a, b), "ranks must hold bit-identical gradients after the exchange"
