#!/usr/bin/env python
r"""One guided score evaluation (U-Net forward + input-gradient over all windows) of the bench workload, for
kernel-level ncu captures:

    ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
        -k regex:'patch_kernel<2, 1, 2' --launch-skip 1 -c 1 -o gpurun_out/ln2_c96 python tools/one_eval.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch

import bench
import sda_b200.score as sc


def main():
    length = int(sys.argv[1]) if len(sys.argv) > 1 else bench.LENGTH
    device = torch.device('cuda', 0)
    score = bench.make_score(bench.SIZE, device)
    x, y = bench.synthetic(1, length, bench.SIZE)
    guided = sc.GaussianScore(y.to(device), A=bench.observation, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2).to(device)
    x, t = x.to(device), torch.tensor(0.5, device=device)
    out = guided(x, t)
    torch.cuda.synchronize()
    print('ok', float(out.abs().mean()))
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 0

    if reps:  # timing of `reps` evaluations (e.g. length 12 = the 8 windows a rank of 8 GPUs holds)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()

        for _ in range(reps):
            out = guided(x, t)

        e1.record()
        torch.cuda.synchronize()
        print(f'{{"length": {length}, "windows": {length - 4}, "ms_per_evaluation": {e0.elapsed_time(e1) / reps:.4f}}}')


if __name__ == '__main__':
    main()
