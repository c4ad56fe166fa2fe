#!/usr/bin/env python
r"""Kolmogorov data generation (BASELINE config 4) on this repository's classes -- what
experiments/kolmogorov/generate.py does with one Slurm task per trajectory (`simulate`, :15-26) and one
aggregation job (`aggregate`, :29-53), run as ONE ensemble per GPU:

    for i in seeds:  random.seed(i); x_i = chain.prior()              (generate.py:20-22)
    x = chain.trajectory(stack(x_i), length=128)[64:]                 (:23-24; all members in one library call)
    x = KolmogorovFlow.coarsen(x, 4)                                   (:53)
    train / valid / test = 80 / 10 / 10 % of the members              (:34-41)

Outputs `<out>/{train,valid,test}.npy`, arrays (n, 64, 2, size / 4, size / 4) float32 -- the layout of the
reference's HDF5 dataset 'x', readable by sda.utils.TrajectoryDataset (h5py is not in this image).
Under torchrun every rank simulates its contiguous share of the seeds (no collective) and rank 0
concatenates the shards.

    python tools/generate_kolmogorov.py --out /tmp/kolmogorov --members 1024 [--size 256] [--length 128]
"""
import argparse
import os
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np
import torch


def simulate(chain, seeds, length: int, keep: int, coarsen: int, chunk: int = 128):
    r"""Members `seeds` -> (len(seeds), keep, 2, size / coarsen, size / coarsen) on the host."""

    from sda_b200.mcs import KolmogorovFlow

    out = []

    for s0 in range(0, len(seeds), chunk):
        priors = []

        for i in seeds[s0:s0 + chunk]:
            random.seed(i)
            priors.append(chain.prior())

        x = torch.stack(priors).cuda()
        x = chain.trajectory(x, length=length)[length - keep:]      # (keep, E, 2, N, N), on the GPU
        x = KolmogorovFlow.coarsen(x, coarsen) if coarsen > 1 else x
        out.append(x.transpose(0, 1).cpu())

    return torch.cat(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', type=Path, required=True)
    ap.add_argument('--members', type=int, default=1024)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--dt', type=float, default=0.2)
    ap.add_argument('--length', type=int, default=128)
    ap.add_argument('--keep', type=int, default=64)
    ap.add_argument('--coarsen', type=int, default=4)
    args = ap.parse_args()

    from sda_b200.mcs import KolmogorovFlow

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    args.out.mkdir(parents=True, exist_ok=True)
    per = -(-args.members // world)
    seeds = list(range(rank * per, min(args.members, (rank + 1) * per)))
    chain = KolmogorovFlow(size=args.size, dt=args.dt)
    t0 = time.perf_counter()
    x = simulate(chain, seeds, args.length, args.keep, args.coarsen)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    np.save(args.out / f'shard_{rank:03d}.npy', x.numpy())
    print(f'rank {rank}: {len(seeds)} members x {args.length} transitions x {chain.steps} inner steps in {sec:.2f} s '
          f'({len(seeds) * args.length * chain.steps / sec:.0f} member-inner-steps/s incl. host copies)')

    if world > 1:
        import torch.distributed as dist

        dist.init_process_group('gloo')
        dist.barrier()

    wall = time.perf_counter() - t0  # slowest rank: simulation, device -> host copies and the shard file

    if rank == 0:
        x = np.concatenate([np.load(args.out / f'shard_{r:03d}.npy') for r in range(world)])
        i, j = int(0.8 * len(x)), int(0.9 * len(x))

        for name, part in (('train', x[:i]), ('valid', x[i:j]), ('test', x[j:])):
            np.save(args.out / f'{name}.npy', part)

        for r in range(world):
            (args.out / f'shard_{r:03d}.npy').unlink()

        print(f'wrote {args.out}/{{train,valid,test}}.npy: {i} / {j - i} / {len(x) - j} trajectories of shape {x.shape[1:]}')
        import json

        print(json.dumps({
            'workload': f'BASELINE config 4: KolmogorovFlow(size={args.size}, dt={args.dt}), {args.members} members x {args.length} '
                        f'transitions x {chain.steps} inner steps, last {args.keep} kept, coarsened x{args.coarsen}',
            'n_gpus': world, 'members_per_gpu': per, 'wall_s_simulate_and_write_shards': wall,
            'member_inner_steps_per_s_all_gpus': args.members * args.length * chain.steps / wall,
            'member_transitions_per_s_all_gpus': args.members * args.length / wall,
        }))


if __name__ == '__main__':
    main()
