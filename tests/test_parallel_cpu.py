r"""Window sharding (sda_b200.parallel) on CPU: world_size 2 over gloo, MLP kernel (the sharding
logic is independent of which kernel evaluates the windows).  Checks that the sharded score and the
sharded guided score (backward through the all-gather) equal the unsharded ones, and that all ranks
hold bit-identical results."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sda_b200.parallel import window_range


def test_window_range_partitions_everything():
    for n in (1, 7, 60, 61, 120):
        for world in (1, 2, 3, 8):
            spans = [window_range(n, r, world) for r in range(world)]
            per = spans[0][2]
            assert all(s[2] == per for s in spans) and per * world >= n
            covered = [i for b, e, _ in spans for i in range(b, e)]
            assert covered == list(range(n))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, L, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)

    try:
        import sda_b200.score as sc
        from sda_b200.parallel import shard_windows

        torch.manual_seed(0)
        score = sc.MCScoreNet(3, order=2, embedding=16, hidden_features=[32, 32], activation=torch.nn.SiLU)
        x = torch.randn(2, L, 3)
        t = torch.tensor(0.4)
        y = torch.randn(2, (L + 3) // 4, 1)
        A = lambda v: v[..., ::4, :1]  # noqa: E731

        def guided(s):
            return sc.GaussianScore(y, A=A, std=0.05, sde=sc.VPSDE(s, shape=()), gamma=3e-2)(x, t)

        with torch.no_grad():
            plain = score(x, t)
        plain_g = guided(score)

        shard_windows(score)

        with torch.no_grad():
            sharded = score(x, t)
        sharded_g = guided(score)

        # CPU GEMMs block differently for different row counts, so sharded vs unsharded agree to
        # rounding here; the CUDA kernels are batch-invariant and tests/test_gpu_parallel.py checks
        # bit equality there.
        ok = torch.allclose(plain, sharded, rtol=1e-5, atol=1e-6) and torch.allclose(plain_g, sharded_g, rtol=1e-4, atol=1e-5)
        # every rank must hold bit-identical tensors
        gathered = [torch.empty_like(sharded_g) for _ in range(world)]
        dist.all_gather(gathered, sharded_g)
        ok = ok and all(torch.equal(g, sharded_g) for g in gathered)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('L', [9, 12])  # 5 and 8 windows per batch element: uneven and even splits
def test_sharded_score_equals_unsharded(L):
    world = 2
    ctx = mp.get_context('spawn')
    results = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, results)) for r in range(world)]

    for p in procs:
        p.start()

    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    assert dict(results) == {0: True, 1: True}


def _grad_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)

    try:
        import sda_b200.score as sc
        from sda_b200.parallel import allreduce_gradients

        torch.manual_seed(0)
        score = sc.MCScoreNet(3, order=1, embedding=16, hidden_features=[32, 32], activation=torch.nn.SiLU)
        sde = sc.VPSDE(score.kernel, shape=(9,))
        torch.manual_seed(10 + rank)  # every rank its own batch, noise and times
        sde.loss(torch.randn(16, 9)).backward()
        local = [p.grad.clone() for p in sde.parameters()]
        allreduce_gradients(sde)
        ok = True

        for p, g in zip(sde.parameters(), local):
            both = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(both, g)
            ok = ok and torch.allclose(p.grad, sum(both) / world, rtol=1e-6, atol=1e-7)

        # sampling after a sharded setup: ranks seeded differently draw the same trajectory (shard-root broadcast)
        from sda_b200.parallel import shard_windows

        shard_windows(score)
        torch.manual_seed(100 + rank)
        sample = sc.VPSDE(score, shape=(7, 3)).sample((2,), steps=3, corrections=1, tau=0.5)
        both = [torch.empty_like(sample) for _ in range(world)]
        dist.all_gather(both, sample)
        ok = ok and torch.isfinite(sample).all().item() and torch.allclose(both[0], both[1], rtol=1e-5, atol=1e-6)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_seed_broadcast():
    r"""allreduce_gradients averages every parameter gradient over the ranks (the data-parallel training step);
    a window-sharded sampler starts from rank 0's noise whatever the ranks' own generators hold."""

    world = 2
    ctx = mp.get_context('spawn')
    results = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, results)) for r in range(world)]

    for p in procs:
        p.start()

    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    assert dict(results) == {0: True, 1: True}


def test_shard_windows_rejects_unknown_transport():
    r"""`shard_windows(transport=...)`: 'peer' (default, NVLink peer-memory kernel on the fused CUDA path) or 'nccl'."""

    import sda_b200.score as sc
    from sda_b200.parallel import shard_windows

    score = sc.MCScoreNet(2, order=1, hidden_features=[8], activation=torch.nn.SiLU)

    with pytest.raises(ValueError):
        shard_windows(score, transport='carrier-pigeon')

    assert shard_windows(score, transport='nccl') is score and score._sdab_sharded
