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
    out = guided(x.to(device), torch.tensor(0.5, device=device))
    torch.cuda.synchronize()
    print('ok', float(out.abs().mean()))


if __name__ == '__main__':
    main()
