"""world_size-2 gloo test (CPU) of the data-parallel host logic: flat gradient all-reduce and the
initial state broadcast.  The model math itself is GPU-only; here plain tensors stand in for it."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointnetgpd_b200.ddp import FlatGradAllReduce, broadcast_module_state


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
        broadcast_module_state(net, src=0)
        w0 = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        gathered = [torch.zeros_like(w0) for _ in range(world)]
        dist.all_gather(gathered, w0)
        same_start = all(torch.equal(gathered[0], g) for g in gathered)

        sync = FlatGradAllReduce(list(net.parameters()), world)
        x = torch.randn(6, 7)          # different data per rank (seeded by rank)
        net(x).square().sum().backward()
        local = [p.grad.clone() for p in net.parameters()]
        l_ptrs = {id(p): p.grad.data_ptr() for p in net.parameters()}
        sync.all_reduce()
        # expected: mean over ranks of the local gradients
        ok = True
        for p, l in zip(net.parameters(), local):
            buf = [torch.zeros_like(l) for _ in range(world)]
            dist.all_gather(buf, l)
            exp = sum(buf) / world
            ok = ok and torch.allclose(p.grad, exp, atol=1e-6)
            ok = ok and p.grad.data_ptr() == l_ptrs[id(p)]              # averaged in place: gradient tensors keep their storage
        # a second step must work with p.grad reset to None (zero_grad(set_to_none=True))
        for p in net.parameters():
            p.grad = None
        net(x).square().sum().backward()
        sync.all_reduce()
        q.put((rank, same_start, ok, sync.bytes_per_step))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_start, ok, nbytes in res:
        assert same_start and ok
        assert nbytes == (7 * 5 + 5 + 5 + 5 + 5 * 3 + 3) * 4


def test_single_process_is_noop():
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    s = FlatGradAllReduce([p], world_size=1)
    s.all_reduce()
    assert torch.equal(p.grad, torch.full((3,), 2.0))
