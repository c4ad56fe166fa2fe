#!/usr/bin/env python
r"""Headline benchmark: denoising steps/sec of guided posterior sampling on synthetic Kolmogorov
256 x 256, L = 64 trajectories (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One denoising step = one iteration of the loop of VPSDE.sample (reference sda/score.py:250-261) =
1 predictor + 1 corrector update = 2 guided score evaluations, each a U-Net forward AND
input-gradient over all 60 trajectory windows (GaussianScore, score.py:375-396).
Network = experiments/kolmogorov/train.py CONFIG (window 5, channels 96/192/384, blocks 3/3/3,
SiLU, circular padding) at 256 x 256 with random-init weights; observation = every 4th frame
coarsened x8 (experiments/kolmogorov/figures.ipynb:204), std 0.1, gamma 1e-2, tau 0.5.

Multi-GPU: the window batch is sharded over the ranks (strong scaling: the trajectory is fixed), one
all-gather of window scores (and of window input-gradients in the backward) per evaluation.

`--impl reference` times the CPU restatement of the reference algorithm (oracle/score_oracle.py, plain
PyTorch ATen ops with all host threads -- the reference itself is pure PyTorch and does not travel
to the GPU box) on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault('SDAB_NO_TQDM', '1')

import torch  # noqa: E402

WINDOW, CHANNELS, BLOCKS = 5, (96, 192, 384), (3, 3, 3)
SIZE, LENGTH = 256, 64
SCHEDULE_STEPS, CORRECTIONS, TAU = 256, 1, 0.5
VARIANTS = {'guided': 'guided (GaussianScore, detach=False)', 'detach': 'guided without back-propagation (GaussianScore, detach=True)',
            'unguided': 'unguided (eps = MCScoreNet)'}
CONV_FLOP_PER_PIXEL = 6_837_696  # forward conv FLOPs per output pixel and window


# ------------------------------------------------------------------------------------ model
def make_score(size: int, device):
    r"""make_score / LocalScoreUNet of experiments/kolmogorov/utils.py:29-70 on this repo's classes."""

    import sda_b200.score as sc

    class LocalScoreUNet(sc.ScoreUNet):
        def __init__(self, channels, size=64, **kwargs):
            super().__init__(channels, 1, **kwargs)
            domain = 2 * torch.pi / size * (torch.arange(size) + 1 / 2)
            self.register_buffer('forcing', torch.sin(4 * domain).expand(1, size, size).clone())

        def forward(self, x, t, c=None):
            return super().forward(x, t, self.forcing)

    torch.manual_seed(0)
    score = sc.MCScoreNet(2, order=WINDOW // 2)
    score.kernel = LocalScoreUNet(
        WINDOW * 2, size, embedding=64, hidden_channels=CHANNELS, hidden_blocks=BLOCKS, kernel_size=3,
        activation=torch.nn.SiLU, spatial=2, padding_mode='circular',
    )

    return score.to(device)


def observation(v):
    r"""A(x): every 4th frame, coarsened x8 (figures.ipynb:204)."""

    from sda_b200.mcs import KolmogorovFlow

    return KolmogorovFlow.coarsen(v[:, ::4], 8)


def synthetic(batch: int, length: int, size: int):
    g = torch.Generator().manual_seed(0)
    x = torch.randn((batch, length, 2, size, size), generator=g)
    y = torch.randn((batch, (length + 3) // 4, 2, size // 8, size // 8), generator=g)
    return x, y


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(',')])

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}

        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 9]

        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}

        sm = sorted(float(r[1]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for j, n in enumerate(names) if any(r[5 + j].lower().startswith('active') for r in rows)]

        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(rows[0][2]), 'reasons': reasons, 'samples': len(rows),
                'power_w_max': max(float(r[3]) for r in rows)}


# ------------------------------------------------------------------------------------ CPU baseline
CPU_WINDOWS = 2  # windows of the bounded CPU sample (L = 6); BASELINE.md section 5 plans a 2-window slice


def cpu_step_seconds(n_steps: int, warm: int, windows: int = CPU_WINDOWS):
    r"""Seconds per denoising step of the oracle (reference algorithm, torch CPU, all host threads) on a
    `windows`-window slice (L = windows + 4) of the 256 x 256 workload.  Returns (seconds per sample
    step, cores)."""

    from oracle import score_oracle as so

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    score = make_score(SIZE, 'cpu')
    state = {k[len('kernel.'):]: v for k, v in score.state_dict().items()}
    k = WINDOW // 2
    x, y = synthetic(1, windows + 2 * k, SIZE)
    eps_fn = lambda a, b: so.gaussian_score(lambda c, d: so.mc_score(state, c, d, k), y, observation, 0.1, a, b, gamma=1e-2)  # noqa: E731
    g = torch.Generator().manual_seed(1)
    noise = [torch.randn(x.shape, generator=g) for _ in range((n_steps + warm) * CORRECTIONS)]
    times = []

    for i in range(n_steps + warm):
        t0 = time.perf_counter()
        # one loop iteration at schedule position i (cost does not depend on i)
        so.pc_sample(eps_fn, x, steps=SCHEDULE_STEPS, corrections=CORRECTIONS, tau=TAU,
                     noise=noise[i * CORRECTIONS:(i + 1) * CORRECTIONS], n_steps=1)
        times.append(time.perf_counter() - t0)

    timed = times[warm:]

    return sum(timed) / len(timed), cores


def gpu_eager_baseline(device) -> dict:
    r"""The "kernel to beat" on the same box: the reference
    algorithm as stock PyTorch-eager CUDA ops (the oracle port moved to the GPU: F.pad(circular) + cuDNN conv2d,
    ATen LayerNorm chain, autograd for the input-gradient), same network, inputs and guided step, with cuDNN /
    cuBLAS TF32 allowed (PyTorch's default for convolutions) and disallowed.  The autograd state of 60 windows
    at 256 x 256 does not fit 180 GB, and windows are independent, so the step is timed on a slice of `nw`
    windows and scaled linearly.  Also reports each mode's guided-score rel-L2 against the fp32 CPU oracle on a
    2-window slice: the parity bar the repo's own path is held to is 1e-4."""

    from oracle import score_oracle as so
    from oracle.testing import rel_l2

    k = WINDOW // 2
    windows = LENGTH - 2 * k
    score = make_score(SIZE, 'cpu')
    state_cpu = {n[len('kernel.'):]: v for n, v in score.state_dict().items()}
    state = {n: v.to(device) for n, v in state_cpu.items()}
    flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {'kind': 'reference algorithm (oracle port), PyTorch eager on the same GPU', 'unit': 'steps/s'}

    # parity sample: L = 6 (2 windows) guided score, CPU fp32 oracle as the yardstick
    xs, ys = synthetic(1, 2 + 2 * k, SIZE)
    ts = torch.tensor(0.5)
    ref = so.gaussian_score(lambda a, b: so.mc_score(state_cpu, a, b, k), ys, observation, 0.1, xs, ts, gamma=1e-2)

    try:
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            name = 'tf32' if tf32 else 'fp32'
            got = so.gaussian_score(lambda a, b: so.mc_score(state, a, b, k), ys.to(device), observation, 0.1,
                                    xs.to(device), ts.to(device), gamma=1e-2)
            out[f'{name}_rel_l2_vs_cpu_fp32'] = rel_l2(got, ref)
            del got

            for nw in (12, 6, 3):
                try:
                    x, y = synthetic(1, nw + 2 * k, SIZE)
                    x, y = x.to(device), y.to(device)
                    eps_fn = lambda a, b: so.gaussian_score(lambda c, d: so.mc_score(state, c, d, k), y, observation, 0.1, a, b, gamma=1e-2)  # noqa: E731
                    g = torch.Generator().manual_seed(1)
                    noise = [torch.randn(x.shape, generator=g).to(device) for _ in range(3)]
                    so.pc_sample(eps_fn, x, steps=SCHEDULE_STEPS, corrections=CORRECTIONS, tau=TAU, noise=noise[:1], n_steps=1)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(2):
                        so.pc_sample(eps_fn, x, steps=SCHEDULE_STEPS, corrections=CORRECTIONS, tau=TAU, noise=noise[1 + i:2 + i], n_steps=1)
                    e1.record()
                    torch.cuda.synchronize()
                    sec = e0.elapsed_time(e1) * 1e-3 / 2
                    out[name] = 1.0 / (sec * windows / nw)
                    out[f'{name}_sample'] = f'2 denoising steps on a {nw}-window slice (L={nw + 4}), {sec * 1e3:.0f} ms each, scaled x{windows / nw:g}'
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = flags

    out['flags'] = {'cudnn.allow_tf32 / cuda.matmul.allow_tf32': 'True for the tf32 entry, False for the fp32 entry',
                    'cudnn.benchmark': torch.backends.cudnn.benchmark, 'torch': torch.__version__,
                    'cudnn': torch.backends.cudnn.version()}
    torch.cuda.empty_cache()

    return out


def secondary(device) -> dict:
    r"""The two other measured paths, timed by the same run so that the driver sees them (not the headline):
    the Kolmogorov stepper (BASELINE config 4 per-GPU share: 128 members, 256 x 256, dt = 0.2) against the HBM
    roofline of its streaming formulation (16 N^2 bytes per member and inner step), and
    one training iteration (BASELINE config 5 per-GPU share: batch 32 at 64 x 64, VPSDE.loss + backward +
    AdamW) against the tensor roofline (3 x forward FLOPs)."""

    import sda_b200.score as sc
    from sda_b200.mcs import KolmogorovFlow

    peaks_file = ROOT / 'MEASURED_PEAKS.json'
    peaks = json.loads(peaks_file.read_text()) if peaks_file.exists() else {}
    hbm, tflops = float(peaks.get('hbm_gbs', 6550.0)), float(peaks.get('bf16_tflops_sustained', 1400.0))
    out = {}

    size, E, transitions = 256, 128, 4
    chain = KolmogorovFlow(size=size, dt=0.2)
    x = chain.prior((E,)).to(device)
    chain.trajectory(x, 1, last=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = chain.trajectory(x, transitions, last=True)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    rate = transitions * chain.steps * E / sec
    gbs = rate * 16 * size * size / 1e9
    out['stepper'] = {
        'workload': f'KolmogorovFlow(size={size}, dt=0.2), {E} members, {transitions} transitions x {chain.steps} inner steps',
        'member_inner_steps_per_s': rate, 'member_transitions_per_s': transitions * E / sec, 'finite': bool(torch.isfinite(y).all()),
        'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm, 'traffic': None,
                     'note': 'algorithmic bytes = 16 N^2 per member and inner step (read + write u, v once)'},
    }
    del x, y, chain
    torch.cuda.empty_cache()

    score = make_score(64, device)
    sde = sc.VPSDE(score.kernel, shape=(10, 64, 64)).to(device).train()
    from sda_b200.parallel import PeerAdamW

    opt = PeerAdamW(sde, lr=2e-4, weight_decay=1e-3)  # one GPU here: the fused AdamW kernel of csrc/peer.cu
    xb = torch.randn(32, 10, 64, 64, device=device, generator=torch.Generator(device=device).manual_seed(0))

    def it():
        l = sde.loss(xb)
        opt.zero_grad(set_to_none=True)
        l.backward()
        opt.step()

    for _ in range(3):
        it()

    torch.cuda.synchronize()
    iters = 8
    e0.record()
    for _ in range(iters):
        it()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / iters
    tf = 3 * 32 * CONV_FLOP_PER_PIXEL * 64 * 64 / sec / 1e12
    out['training'] = {
        'workload': 'VPSDE.loss + backward + AdamW (sda_b200.parallel.PeerAdamW), windows (10, 64, 64), batch 32, mode ' + os.environ.get('SDAB_MODE', 'bf16x3'),
        'ms_per_iteration': sec * 1e3, 'samples_per_s': 32 / sec,
        'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': tflops, 'unit': 'TFLOP/s', 'frac': tf / tflops, 'traffic': None,
                     'note': 'algorithmic FLOPs = 3 x forward convolution FLOPs (forward, input-gradient, weight-gradient)'},
    }

    return out


def run_reference(args, rank):
    if rank != 0:
        return

    windows = LENGTH - 2 * (WINDOW // 2)
    nw = windows if args.full_cpu else CPU_WINDOWS
    sec, cores = cpu_step_seconds(args.steps, args.warmup, nw)
    value = 1.0 / (sec * windows / nw)
    sample = (f'one denoising step (2 guided score evaluations, U-Net forward + input-gradient) on a {nw}-window slice '
              f'(L={nw + 4}) of the {SIZE}x{SIZE} workload, {sec:.2f} s per sample step, scaled x{windows / nw:g}')
    print(json.dumps({
        'impl': 'reference',
        'metric': 'denoising steps/sec, Kolmogorov 256x256 L=64 guided posterior sampling',
        'value': value, 'unit': 'steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 / value, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def workload_config(args, world):
    return {
        'workload': f'Kolmogorov {SIZE}x{SIZE}, L={LENGTH}, B={args.batch}, window {WINDOW} (k=2) -> {args.batch * (LENGTH - 4)} windows; '
                    f'U-Net {CHANNELS} x {BLOCKS}; {VARIANTS[args.variant]}, corrections={CORRECTIONS}, tau={TAU}',
        'windows': args.batch * (LENGTH - 4),
        'score_evaluations_per_step': 1 + CORRECTIONS,
        'parallelism': f'window-sharded x{world}' if world > 1 else 'single GPU',
        'mode': os.environ.get('SDAB_MODE', 'bf16x3'),
        'l2': 'per-step working set (tens of GB of activations) exceeds the 126 MB L2; no explicit flush needed',
    }


def exchange_used(score, world):
    r"""How the window shards were exchanged in this run (what ran, not what was asked for)."""

    if world == 1:
        return {}

    import sda_b200.score as sc

    bufs = list(score.kernel.network._buffers_mc.values())
    peer = [b for b in bufs if isinstance(b, sc.PeerExchange)]

    return {'exchange': 'peer-memory all-gather kernel over NVLink (csrc/peer.cu), one launch per exchange' if peer and len(peer) == len(bufs)
            else 'NCCL all_gather_into_tensor'}


# ------------------------------------------------------------------------------------ ours
def run_ours(args, rank, local_rank, world):
    import ctypes

    import torch.distributed as dist

    import sda_b200.score as sc
    from sda_b200 import _lib
    from sda_b200.parallel import shard_windows

    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)

    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    lib = _lib.load()
    score = make_score(SIZE, device)

    if world > 1:
        shard_windows(score, transport=args.exchange)

    x_host, y_host = synthetic(args.batch, LENGTH, SIZE)
    x_pin, y_pin = x_host.pin_memory(), y_host.pin_memory()
    # --variant: guided (the metric: GaussianScore, forward + input-gradient), detach (GaussianScore(detach=True),
    # score.py:378-379: no back-propagation through the network) or unguided (eps = MCScoreNet, prior sampling)
    guided = None

    if args.variant == 'unguided':
        sde = sc.VPSDE(score, shape=tuple(x_host.shape[1:])).to(device)
    else:
        guided = sc.GaussianScore(y_host.to(device), A=observation, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2,
                                  detach=args.variant == 'detach').to(device)
        sde = sc.VPSDE(guided, shape=tuple(x_host.shape[1:])).to(device)
    x = x_host.to(device)
    state = sde.sampler_state(x, SCHEDULE_STEPS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        tv = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return float(tv)

    step = 0

    for _ in range(args.warmup):
        x = sde.denoise_step(x, step, state, corrections=CORRECTIONS, tau=TAU)
        step += 1

    # ---------------- device-resident timed region
    sampler = ClockSampler(local_rank)

    if rank == 0:
        sampler.start()

    barrier()
    _lib.launch_count(reset=True)
    lib.sdab_conv_profile(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()

    for _ in range(args.steps):
        x = sde.denoise_step(x, step, state, corrections=CORRECTIONS, tau=TAU)
        step += 1

    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count()
    conv_ms, conv_flops, conv_launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    _lib.check(lib.sdab_conv_profile_read(ctypes.byref(conv_ms), ctypes.byref(conv_flops), ctypes.byref(conv_launches)))
    lib.sdab_conv_profile(0)
    clocks = sampler.stop() if rank == 0 else None
    assert os.environ.get('SDAB_UMMA_DEBUG') or torch.isfinite(x).all(), 'non-finite state after the timed steps'
    # proof carried by the line itself: the state after warmup + steps denoising steps.  The sharded path is
    # bit-identical to the single-GPU one, so this hash is the same at N = 1, 2, 4, 8 for equal --steps / --warmup
    import hashlib

    state_sha = hashlib.sha256(x.cpu().numpy().tobytes()).hexdigest()[:16]

    # ---------------- end to end through the public step API, host buffers in and out
    barrier()
    t0 = time.perf_counter()

    for _ in range(0 if args.profile_run else args.steps):
        x.copy_(x_pin, non_blocking=True)
        if guided is not None:
            guided.y.copy_(y_pin, non_blocking=True)
        x = sde.denoise_step(x, step % SCHEDULE_STEPS, state, corrections=CORRECTIONS, tau=TAU)
        x_pin.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
        step += 1

    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    if args.trace is not None:
        # kernel-level timeline of two more steps (after everything that is timed): where the non-convolution
        # time of a step goes (NCCL, window maps, sampler kernels, the user's A), per kernel name
        from torch.profiler import ProfilerActivity, profile

        barrier()

        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(2):
                x = sde.denoise_step(x, step % SCHEDULE_STEPS, state, corrections=CORRECTIONS, tau=TAU)
                step += 1
            torch.cuda.synchronize()

        if rank == 0:
            args.trace.parent.mkdir(parents=True, exist_ok=True)
            rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
            rows.sort(key=lambda r: -r[1])
            total = sum(r[1] for r in rows)
            with open(args.trace, 'w') as f:
                f.write(f'# torch.profiler, 2 denoising steps on rank 0 of {world}: device time per kernel (ms), total {total:.2f} ms\n')
                for name, ms, count in rows:
                    f.write(f'{ms:10.3f} ms {100 * ms / total:5.1f}% x{count:5d}  {name[:140]}\n')

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_file = ROOT / 'MEASURED_PEAKS.json'
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
    peak = 1400.0

    if peaks_file.exists():
        peak = float(json.loads(peaks_file.read_text()).get('bf16_tflops_sustained', peak))
    else:
        peak_src = 'fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)'

    achieved = conv_flops.value / (conv_ms.value * 1e-3) / 1e12 if conv_ms.value > 0 else 0.0
    passes = 3 if os.environ.get('SDAB_MODE', 'bf16x3') == 'bf16x3' else 1
    traffic = None
    traffic_file = ROOT / 'profiles' / 'conv_umma_traffic.json'

    if traffic_file.exists():
        traffic = json.loads(traffic_file.read_text()).get('dram_bytes_per_launch')

    value = args.steps / (ms * 1e-3)
    out = {
        'metric': 'denoising steps/sec, Kolmogorov 256x256 L=64 guided posterior sampling',
        'value': value, 'unit': 'steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'bf16x3 (split-bf16 operands, fp32 accumulate)' if passes == 3 else 'bf16 (fp32 accumulate)',
        'data': 'synthetic', 'config': dict(workload_config(args, world), **exchange_used(score, world)), 'clocks': clocks,
        'e2e': {'value': None if args.profile_run else args.steps / e2e_s, 'unit': 'steps/s', 'h2d_bytes_per_step': x_pin.numel() * 4 + y_pin.numel() * 4,
                'd2h_bytes_per_step': x_pin.numel() * 4},
        'gpu_launches': int(launches), 'state_sha': state_sha,
        'roofline': {
            'bound': 'tensor', 'kernel': 'conv_umma_patch_kernel (+ conv_umma_kernel for strided / sub-pixel layers)', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
            'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
            'traffic_source': 'profiles/conv_umma_traffic.json: DRAM bytes per convolution launch from the ncu launch list of the same command '
                              '(not re-measured in this run: ncu cannot run inside a timed bench)',
            'note': f'algorithmic FLOPs (one multiply-add pair per product) over CUDA-event kernel time of {conv_launches.value} '
                    f'launches on rank 0; the {passes}-pass mode issues {passes}x that many tensor-core FLOPs '
                    f'(tensor-pipe fraction ~ {passes * achieved / peak:.3f})',
            'conv_share_of_step': conv_ms.value / ms,
        },
    }

    if args.profile_run:
        args.no_cpu_baseline = args.no_secondary = True
        out['profile_run'] = 'under a profiler: not a bench value'

    if world == 1 and not args.no_cpu_baseline and args.variant == 'guided':
        sec, cores = cpu_step_seconds(2, 1)
        windows = LENGTH - 2 * (WINDOW // 2)
        out['cpu_baseline'] = {
            'value': 1.0 / (sec * windows * args.batch / CPU_WINDOWS), 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
            'sample': f'2 denoising steps (after 1 warm-up) of the torch-CPU oracle on a {CPU_WINDOWS}-window slice '
                      f'(L={CPU_WINDOWS + 4}) at {SIZE}x{SIZE}, {sec:.1f} s each, scaled x{windows * args.batch / CPU_WINDOWS:g}',
        }
        out['gpu_eager_baseline'] = gpu_eager_baseline(device)

    if world == 1 and not args.no_secondary and args.variant == 'guided':
        out['secondary'] = secondary(device)

    print(json.dumps(out))

    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=1, help='trajectories sampled together (B)')
    ap.add_argument('--exchange', default='peer', choices=['peer', 'nccl'],
                    help='N > 1: exchange of the window shards (peer-memory kernel over NVLink, or NCCL all-gather)')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the cpu_baseline and gpu_eager_baseline legs')
    ap.add_argument('--no-secondary', action='store_true', help='skip the stepper / training measurements')
    ap.add_argument('--profile-run', action='store_true', help='for runs under ncu: one warm-up step allowed, no e2e / baseline legs (not a bench value)')
    ap.add_argument('--trace', type=Path, default=None, help='write a torch.profiler kernel table of two extra steps (rank 0) to this file')
    ap.add_argument('--full-cpu', action='store_true', help='--impl reference: time all 60 windows (about 90 s per step)')
    ap.add_argument('--variant', default='guided', choices=sorted(VARIANTS),
                    help='guided is the BASELINE metric; the others are reported beside it (SURVEY.md section 8d)')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    args.warmup = max(args.warmup, 1 if args.profile_run else 3) if args.impl == 'ours' else args.warmup

    if args.impl == 'reference':
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
