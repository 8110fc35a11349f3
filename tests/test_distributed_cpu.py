"""N > 1 host logic on CPU: gloo, world_size 2.  The solve itself is injected (the CPU oracle) so
that sharding, the weight broadcast and the gather can be checked without a GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchcde_b200 import distributed as D


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 8, 65536, 524288 + 3):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import math
    from oracle import cde_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different initial weights on every rank
    lin = torch.nn.Linear(4, 4 * 3)
    gen = torch.Generator().manual_seed(0)              # same data everywhere, sharded below
    n, length = 11, 9
    x = torch.randn(n, length, 3, generator=gen).cumsum(1) / math.sqrt(length)
    z0 = torch.randn(n, 4, generator=gen)
    lo, hi = D.shard_bounds(n, world, rank)
    t = torch.tensor([0.0, length - 1.0])

    def solve(X, func, z, tt, **kw):
        return O.cdeint_linear(X, O.knot_times(length, torch.float32), func.weight, func.bias, z, tt, "rk4", 1.0)

    out = D.cdeint_sharded(solve, lambda: O.hermite_backward_difference_coeffs(x[lo:hi]), lin, z0[lo:hi], t,
                           gather=True)
    if rank == 0:
        full = O.cdeint_linear(O.hermite_backward_difference_coeffs(x), O.knot_times(length, torch.float32),
                               lin.weight, lin.bias, z0, t, "rk4", 1.0)
        q.put((bool(torch.equal(out, full)), lin.weight.detach().flatten().tolist()))
    else:
        q.put((True, lin.weight.detach().flatten().tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_matches_single_process_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _ in got)
    assert got[0][1] == got[1][1]            # the broadcast made the weights identical
