r"""Window sharding over 2 GPUs (NCCL): sharded score and guided score are BIT-identical to the
single-GPU ones (the CUDA kernels are batch-invariant and the overlap-add runs in fixed order)."""

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, results):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))

    try:
        import sys
        from pathlib import Path

        sys.path.insert(0, str(Path(__file__).resolve().parent))
        import sda_b200.score as sc
        from helpers import build_score
        from oracle.testing import randn
        from sda_b200.parallel import shard_windows

        ok = True

        for transport, (B, L) in ((tr, bl) for tr in ('peer', 'nccl') for bl in ((1, 9), (2, 6))):
            # 7 windows: uneven 4 / 3 split; 2 x 4 windows: one trajectory per rank
            score, k = build_score('net_small', 32, 'cuda')
            x = randn((B, L, 2, 32, 32), seed=1).cuda()
            y = randn((B, L, 2, 16, 16), seed=2).cuda()
            t = torch.tensor(0.45).cuda()
            A = lambda v: v[..., ::2, ::2]  # noqa: E731

            def guided(s):
                return sc.GaussianScore(y, A=A, std=0.1, sde=sc.VPSDE(s, shape=()), gamma=1e-2).cuda()(x, t)

            with torch.no_grad():
                plain = score(x, t)
            plain_g = guided(score)
            shard_windows(score, transport=transport)

            for _ in range(3):  # the peer exchange alternates two buffers: several turns of each
                with torch.no_grad():
                    sharded = score(x, t)
                sharded_g = guided(score)
                ok = ok and bool(torch.equal(plain, sharded) and torch.equal(plain_g, sharded_g))

            # the transport that was asked for is the one that ran (no silent fallback to NCCL)
            peers = [v for v in score.kernel.network._buffers_mc.values() if isinstance(v, sc.PeerExchange)]
            ok = ok and (len(peers) == 2 if transport == 'peer' else not peers)
            # the materialised fallback of the sharded path (per-window kernels that are not fusable)
            score.fuse_windows = False
            with torch.no_grad():
                ok = ok and bool(torch.equal(plain, score(x, t)))
            ok = ok and bool(torch.equal(plain_g, guided(score)))
            score.fuse_windows = True

        # ranks seeded differently still sample the same trajectory: rank 0's noise and Philox seed are broadcast
        torch.manual_seed(100 + rank)
        sde = sc.VPSDE(sc.GaussianScore(y, A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2), shape=(L, 2, 32, 32)).cuda()
        sample = sde.sample((B,), steps=2, corrections=1, tau=0.5)
        both = [torch.empty_like(sample) for _ in range(world)]
        dist.all_gather(both, sample)
        ok = ok and bool(torch.isfinite(sample).all()) and all(bool(torch.equal(both[0], o)) for o in both)
        # data-parallel training step: PeerAdamW (reduce + AdamW + broadcast over peer memory) against
        # allreduce_gradients + torch.optim.AdamW on a copy, different data per rank
        import copy

        from sda_b200.parallel import PeerAdamW, allreduce_gradients

        score, k = build_score('net_small', 16, 'cuda')
        sde_a = sc.VPSDE(score.kernel, shape=(6, 16, 16)).cuda().train()
        sde_b = copy.deepcopy(sde_a)
        opt_a = torch.optim.AdamW(sde_a.parameters(), lr=1e-3, weight_decay=1e-2)
        opt_b = PeerAdamW(sde_b, lr=1e-3, weight_decay=1e-2)
        xb = randn((8, 6, 16, 16), seed=10 + rank).cuda()

        for it in range(3):
            for sde_, opt_, avg in ((sde_a, opt_a, True), (sde_b, opt_b, False)):
                torch.manual_seed(7 * it + rank)
                l = sde_.loss(xb)
                opt_.zero_grad()
                l.backward()
                if avg:
                    allreduce_gradients(sde_)
                opt_.step()

        flat = opt_b.params_buf.view.clone()
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        ok = ok and all(bool(torch.equal(both[0], o)) for o in both)  # replicas stay bit-identical
        for pa, pb in zip(sde_a.parameters(), sde_b.parameters()):
            ok = ok and float((pa - pb).norm() / pa.norm().clamp_min(1e-30)) < 2e-5
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_sharding_is_bit_identical():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]

    ctx = mp.get_context('spawn')
    results = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]

    for p in procs:
        p.start()

    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0

    assert dict(results) == {0: True, 1: True}
