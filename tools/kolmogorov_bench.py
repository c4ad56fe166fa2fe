#!/usr/bin/env python
r"""Secondary measurement (BASELINE config 4): Kolmogorov 256 x 256 data generation,
ensemble members per GPU = 1024 / 8 = 128, dt = 0.2 (82 inner steps per transition).

Reports transitions/s, inner steps/s and the effective bandwidth against the algorithmic
16 N^2 bytes per inner step per member, next to the NumPy oracle (the
restated reference algorithm, 1 core like experiments/kolmogorov/generate.py:16) on one member.
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

from oracle import kolmogorov_oracle as ko
from sda_b200 import _lib
from sda_b200.mcs import KolmogorovFlow


def main():
    size, E, transitions = 256, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 8
    chain = KolmogorovFlow(size=size, dt=0.2)
    x = chain.prior((E,)).cuda()
    chain.trajectory(x, 1, last=True)  # warm-up
    torch.cuda.synchronize()
    _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = chain.trajectory(x, transitions, last=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    inner = transitions * chain.steps * E
    out = {
        'workload': f'KolmogorovFlow(size={size}, dt=0.2), ensemble {E}, {transitions} transitions x {chain.steps} inner steps',
        'ms': ms,
        'member_transitions_per_s': transitions * E / (ms * 1e-3),
        'member_inner_steps_per_s': inner / (ms * 1e-3),
        'effective_GBps_vs_16N2': inner * 16 * size * size / (ms * 1e-3) / 1e9,
        'gpu_launches': _lib.launch_count(),
        'finite': bool(torch.isfinite(y).all()),
        'max_speed': float(y.square().sum(dim=1).sqrt().max()),
    }
    # restated reference algorithm on the host, fp32, one member, a few inner steps
    x0 = x[:1].cpu().numpy()
    t0 = time.perf_counter()
    ko.transition(x0, dt=0.2, n_inner=8)
    sec = (time.perf_counter() - t0) / 8
    out['cpu_oracle_inner_steps_per_s_1core'] = 1.0 / sec
    out['speedup_vs_cpu_oracle_1core'] = out['member_inner_steps_per_s'] * sec
    print(json.dumps(out))


if __name__ == '__main__':
    main()
