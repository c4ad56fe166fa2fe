#!/usr/bin/env python
r"""Secondary measurement (BASELINE config 5): score U-Net training on synthetic Kolmogorov 64 x 64
windows, global batch 256 = 32 per GPU, data-parallel (torch DDP over NCCL all-reduces the
22.9 M fp32 gradients).  One iteration = VPSDE.loss (sda/score.py:265-276) forward + backward +
AdamW step, as in sda/utils.py:136-143 with experiments/kolmogorov/train.py's CONFIG.

    python tools/train_bench.py [iterations]
    python -m torch.distributed.run --nproc-per-node N ... tools/train_bench.py [iterations]

Precision: forward and input-gradient run the tensor-core path in SDAB_MODE (bf16x3 by default;
`SDAB_MODE=bf16` is the single-pass mode BASELINE config 5 names), weight gradients run on the
fp32 CUDA cores, master weights and AdamW state are fp32.
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch
import torch.distributed as dist

import bench
from sda_b200 import _lib
import sda_b200.score as sc


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)

    if world > 1:
        dist.init_process_group('nccl')

    score = bench.make_score(64, 'cuda')
    sde = sc.VPSDE(score.kernel, shape=(10, 64, 64)).cuda().train()
    opt = torch.optim.AdamW(sde.parameters(), lr=2e-4, weight_decay=1e-3, fused=True)
    g = torch.Generator(device='cuda').manual_seed(local)
    x = torch.randn(32, 10, 64, 64, device='cuda', generator=g)

    def step():
        l = sde.loss(x)
        opt.zero_grad(set_to_none=True)
        l.backward()
        opt.step()
        return l

    peer = os.environ.get('SDAB_TRAIN_OPT', 'peer') == 'peer' and not os.environ.get('SDAB_TRAIN_DDP')

    if peer:
        # sda_b200.parallel.PeerAdamW: gradient exchange + AdamW + parameter broadcast as one kernel per rank over
        # NVLink peer memory (a plain fused AdamW on one GPU).  SDAB_TRAIN_OPT=torch: torch.optim.AdamW(fused=True)
        # after allreduce_gradients
        from sda_b200.parallel import PeerAdamW

        opt = PeerAdamW(sde, lr=2e-4, weight_decay=1e-3)

        def step():  # noqa: F811
            l = sde.loss(x)
            opt.zero_grad()
            l.backward()
            opt.step()
            return l
    elif world > 1 and os.environ.get('SDAB_TRAIN_DDP'):
        # torch DDP for comparison: its hooks fire on the parameter gradients produced by the native backward;
        # VPSDE.loss is not the module's forward, so route it through a thin wrapper module
        class Loss(torch.nn.Module):
            def __init__(self, sde):
                super().__init__()
                self.sde = sde

            def forward(self, x):
                return self.sde.loss(x)

        wrapped = torch.nn.parallel.DistributedDataParallel(Loss(sde), device_ids=[local])

        def step():  # noqa: F811
            l = wrapped(x)
            opt.zero_grad(set_to_none=True)
            l.backward()
            opt.step()
            return l
    elif world > 1:
        from sda_b200.parallel import allreduce_gradients

        for p in sde.parameters():  # same initial weights on every rank (DDP broadcasts them too)
            dist.broadcast(p.data, src=0)

        def step():  # noqa: F811
            l = sde.loss(x)
            opt.zero_grad(set_to_none=True)
            l.backward()
            allreduce_gradients(sde)  # one in-place all-reduce of the flat convolution-gradient buffer
            opt.step()
            return l

    for _ in range(3):
        step()

    torch.cuda.synchronize()
    _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    if world > 1:
        dist.barrier()

    e0.record()

    for _ in range(iters):
        loss = step()

    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor(e0.elapsed_time(e1), device='cuda')

    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)

    if local == 0:
        ms = float(ms)
        print(json.dumps({
            'workload': 'VPSDE.loss + backward + AdamW, U-Net (96, 192, 384) x (3, 3, 3), windows (10, 64, 64), '
                        f'batch 32 per GPU x {world} GPU(s)',
            'mode': os.environ.get('SDAB_MODE', 'bf16x3'), 'n_gpus': world,
            'optimizer': 'PeerAdamW (csrc/peer.cu)' if peer else 'torch.optim.AdamW(fused=True)',
            'gradient_exchange': 'peer-memory reduce + AdamW + broadcast kernel' if peer and world > 1 else 'none' if world == 1 else ('torch DDP' if os.environ.get('SDAB_TRAIN_DDP') else 'flat in-place all-reduce'), 'iterations': iters,
            'ms_per_iteration': ms / iters, 'iterations_per_s': iters / (ms * 1e-3),
            'samples_per_s': 32 * world * iters / (ms * 1e-3),
            'algorithmic_tflops': 3 * 32 * world * bench.CONV_FLOP_PER_PIXEL * 64 * 64 * iters / (ms * 1e-3) / 1e12,
            'gpu_launches_per_iteration': _lib.launch_count() / iters, 'loss': float(loss.detach()),
        }))

    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
