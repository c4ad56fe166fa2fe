#!/usr/bin/env python
r"""Developer diagnostic: runs every kernel family of libsdab on the GPU against torch / the
oracle and prints one line per check.  Not a test (tests/ holds those); meant for
`gpurun -- python tools/gpu_check.py > gpurun_out/check.log`.

Each check runs in a SUBPROCESS with a timeout so that a hung or trapping kernel
neither stops the remaining checks nor hangs the GPU box.
"""

from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def rel(a, b):
    import torch

    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def conv_ref(x, w, b, stride):
    import torch.nn.functional as F

    return F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode='circular'), w.double(), None if b is None else b.double(), stride=stride)


def run_conv(x, w, b, stride, transpose, mode, engine):
    import torch

    from sda_b200 import _lib

    lib = _lib.load()
    N, _, H, W = x.shape
    Cout, Cin = w.shape[:2]
    nbytes = lib.sdab_conv3x3_workspace_bytes(N, Cin, Cout, H, W, stride, transpose)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device='cuda')
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    co = Cin if transpose else Cout
    out = torch.full((N, co, H // stride, W // stride), float('nan'), device='cuda')
    _lib.check(lib.sdab_conv3x3(x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(), out.data_ptr(), N, Cin, Cout,
                                H, W, stride, transpose, mode, engine, base, nbytes, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out


CONV_CASES = [
    # N, Cin, Cout, H, W, stride
    (2, 32, 32, 16, 16, 1),
    (2, 11, 96, 16, 16, 1),
    (1, 96, 96, 32, 32, 1),
    (3, 96, 96, 4, 4, 1),
    (2, 96, 192, 32, 32, 2),
    (1, 192, 192, 16, 16, 1),
    (1, 384, 384, 16, 16, 1),
    (1, 384, 192, 16, 16, 1),
    (2, 96, 10, 16, 16, 1),
    (1, 96, 96, 8, 256, 1),
    (1, 192, 384, 32, 32, 2),
    (5, 64, 64, 8, 8, 2),
]


def _conv_engine(engine):
    import torch

    torch.manual_seed(0)
    worst = 0.0

    for (N, Cin, Cout, H, W, s) in CONV_CASES:
        x = torch.randn(N, Cin, H, W, device='cuda')
        w = torch.randn(Cout, Cin, 3, 3, device='cuda') / (9 * Cin) ** 0.5
        b = torch.randn(Cout, device='cuda')
        ref = conv_ref(x, w, b, s)

        for mode in (0, 1):
            out = run_conv(x, w, b, s, 0, mode, engine)
            e = rel(out, ref)
            print(f'  conv fwd engine={engine} mode={mode} N={N} {Cin}->{Cout} {H}x{W} s={s}: rel={e:.3e}', flush=True)
            if mode == 0:
                worst = max(worst, e)

        if s == 1:
            g = torch.randn(N, Cout, H, W, device='cuda')
            xr = x.double().requires_grad_(True)
            (gref,) = torch.autograd.grad(conv_ref(xr, w, None, 1), xr, g.double())
            out = run_conv(g, w, None, 1, 1, 0, engine)
            e = rel(out, gref)
            print(f'  conv dgrad engine={engine} N={N} {Cin}<-{Cout} {H}x{W}: rel={e:.3e}', flush=True)
            worst = max(worst, e)

    assert worst < 5e-5, worst


@check
def conv_simt():
    _conv_engine(1)


@check
def conv_umma():
    _conv_engine(0)


@check
def window_maps():
    import numpy as np
    import torch

    from sda_b200.score import MCScoreNet

    g = np.load(ROOT / 'tests/golden/maps.npz')

    for L, k in [(5, 2), (9, 2), (7, 1), (12, 3)]:
        B, C, H, W = 2, 2, 2, 3
        x = torch.arange(B * L * C * H * W, dtype=torch.float32).reshape(B, L, C, H, W).cuda().requires_grad_(True)
        u = MCScoreNet.unfold(x, k)
        assert torch.equal(u.cpu().double(), torch.from_numpy(g[f'unfold_L{L}_k{k}'])), 'unfold'
        f = MCScoreNet.fold(u, k)
        assert torch.equal(f.cpu().double(), torch.from_numpy(g[f'fold_L{L}_k{k}'])), 'fold'
        gg = torch.from_numpy(g[f'adjoint_g_L{L}_k{k}']).float().cuda()
        (gx,) = torch.autograd.grad((u * gg).sum(), x)
        assert torch.equal(gx.cpu().double(), torch.from_numpy(g[f'adjoint_L{L}_k{k}'])), 'unfold adjoint'
        # fold adjoint against torch autograd of the reference formula
        s = torch.randn(B, L - 2 * k, (2 * k + 1) * C, H, W, device='cuda', requires_grad=True)
        go = torch.randn(B, L, C, H, W, device='cuda')
        (a,) = torch.autograd.grad(MCScoreNet.fold(s, k), s, go)
        sr = s.detach().cpu().requires_grad_(True)
        (b,) = torch.autograd.grad(MCScoreNet.fold(sr, k), sr, go.cpu())
        assert torch.equal(a.cpu(), b), 'fold adjoint'
        print(f'  maps L={L} k={k}: exact', flush=True)


@check
def sampler_ops():
    import torch

    from sda_b200 import _lib

    lib = _lib.load()
    B, n = 3, 3 * 1000
    x = torch.randn(n, device='cuda')
    eps = torch.randn(n, device='cuda')
    x0 = x.clone()
    _lib.check(lib.sdab_vpsde_predict(x.data_ptr(), eps.data_ptr(), 0.9, 0.2, n, _lib.stream_ptr()))
    print('  predict', rel(x, 0.9 * x0 + 0.2 * eps))
    assert rel(x, 0.9 * x0 + 0.2 * eps) < 1e-6
    z = torch.randn(n, device='cuda')
    scratch = torch.empty(lib.sdab_vpsde_correct_scratch_floats(B), device='cuda')
    x1 = x.clone()
    _lib.check(lib.sdab_vpsde_correct(x.data_ptr(), eps.data_ptr(), z.data_ptr(), 0.5, 0.7, 0, 0, B, n, scratch.data_ptr(), _lib.stream_ptr()))
    e = eps.view(B, -1)
    delta = 0.5 / e.square().mean(dim=1, keepdim=True)
    ref = x1.view(B, -1) - (delta * e + torch.sqrt(2 * delta) * z.view(B, -1)) * 0.7
    print('  correct', rel(x.view(B, -1), ref))
    assert rel(x.view(B, -1), ref) < 1e-5
    out = torch.empty(1 << 22, device='cuda')
    _lib.check(lib.sdab_randn(out.data_ptr(), out.numel(), 1234, 0, _lib.stream_ptr()))
    m, s = out.mean().item(), out.std().item()
    k = ((out - m) ** 4).mean().item() / s ** 4
    print(f'  randn mean={m:.4f} std={s:.4f} kurt={k:.3f}')
    assert abs(m) < 3e-3 and abs(s - 1) < 3e-3 and abs(k - 3) < 0.05
    out2 = torch.empty(1 << 22, device='cuda')
    _lib.check(lib.sdab_randn(out2.data_ptr(), out2.numel(), 1234, 0, _lib.stream_ptr()))
    assert torch.equal(out, out2)


def _load_net(name, size, device='cuda'):
    import numpy as np
    import torch

    import sda_b200.score as sc
    from oracle.testing import fill_state_

    g = np.load(ROOT / f'tests/golden/{name}.npz')
    cfg = {'net_small': (3, (32, 64), (1, 2), 100), 'net_config': (5, (96, 192, 384), (3, 3, 3), 200)}[name]
    window, ch, blocks, seed = cfg

    class Local(sc.ScoreUNet):
        def __init__(self, channels, size, **kw):
            super().__init__(channels, 1, **kw)
            domain = 2 * torch.pi / size * (torch.arange(size) + 1 / 2)
            self.register_buffer('forcing', torch.sin(4 * domain).expand(1, size, size).clone())

        def forward(self, x, t, c=None):
            return super().forward(x, t, self.forcing)

    score = sc.MCScoreNet(2, order=window // 2)
    score.kernel = Local(window * 2, size, embedding=64, hidden_channels=ch, hidden_blocks=blocks, kernel_size=3,
                         activation=torch.nn.SiLU, spatial=2, padding_mode='circular')
    fill_state_(score.state_dict(), seed=seed)
    return score.to(device), g, window // 2


def _net(name):
    import torch

    import sda_b200.score as sc

    score, g, k = _load_net(name, 16)
    x = torch.from_numpy(g['x']).cuda()
    t = torch.tensor(float(g['t'])).cuda()

    for engine in ('simt', 'umma'):
        os.environ['SDAB_ENGINE'] = engine
        with torch.no_grad():
            eps = score(x, t)
        e32, e64 = rel(eps, torch.from_numpy(g['mc_score'])), rel(eps, torch.from_numpy(g['mc_score_fp64']))
        print(f'  {name} {engine} mc_score: vs ref fp32 {e32:.3e}, vs ref fp64 {e64:.3e}', flush=True)
        A = lambda v: v[..., ::2, ::2]  # noqa: E731
        guided = sc.GaussianScore(torch.from_numpy(g['y']).cuda(), A=A, std=0.1, sde=sc.VPSDE(score, shape=()), gamma=1e-2).cuda()
        gs = guided(x, t)
        eg = rel(gs, torch.from_numpy(g['gaussian_score']))
        print(f'  {name} {engine} gaussian_score: vs ref fp32 {eg:.3e}', flush=True)
        assert e64 < 1e-4 and eg < 1e-4, (e64, eg)


@check
def net_small():
    _net('net_small')


@check
def net_config():
    _net('net_config')


@check
def kolmogorov():
    import numpy as np
    import torch

    from oracle import kolmogorov_oracle as ko
    from sda_b200.mcs import KolmogorovFlow

    for size in (64, 256):
        chain = KolmogorovFlow(size=size, dt=0.2)
        assert chain.steps == ko.inner_steps(size, 0.2)
        rng = np.random.default_rng(0)
        x0 = ko.prior((3,), size, rng, np.float32)
        ref64 = ko.transition(x0.astype(np.float64), dt=0.2)
        out = chain.transition(torch.from_numpy(x0)).numpy()
        e = np.linalg.norm(out - ref64) / np.linalg.norm(ref64)
        div = np.abs(ko.divergence(out.astype(np.float64))).max()
        print(f'  kolmogorov N={size}: one transition rel={e:.3e}, max|div|={div:.3e}', flush=True)
        assert e < 1e-4
        import random

        random.seed(3)
        p = chain.prior((4,))
        speed = p.square().sum(dim=1).sqrt().amax(dim=(-2, -1))
        divp = np.abs(ko.divergence(p.double().numpy())).max()
        print(f'  prior N={size}: max speed {speed.tolist()}, max|div|={divp:.3e}', flush=True)
        assert torch.allclose(speed, torch.full_like(speed, 3.0), atol=1e-3)
        traj = chain.trajectory(p, length=3)
        assert traj.shape == (3, 4, 2, size, size) and torch.isfinite(traj).all()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == '--one':
        CHECKS[sys.argv[2]]()
        return

    names = sys.argv[1:] or list(CHECKS)
    summary = []

    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, '--one', name], capture_output=True, text=True, timeout=300)
            ok = r.returncode == 0
            tail = (r.stdout + r.stderr)[-6000:]
        except subprocess.TimeoutExpired as ex:
            ok = False
            tail = 'TIMEOUT\n' + ((ex.stdout or b'').decode()[-3000:] if isinstance(ex.stdout, bytes) else str(ex.stdout)[-3000:])
        print(f'=== {name}: {"PASS" if ok else "FAIL"} ({time.time() - t0:.1f}s)\n{tail}', flush=True)
        summary.append((name, ok))

    print('SUMMARY', summary)


if __name__ == '__main__':
    main()
