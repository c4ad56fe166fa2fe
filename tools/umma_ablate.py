#!/usr/bin/env python
r"""Developer ablation of conv_umma_kernel: times one convolution with parts of the pipeline
disabled (SDAB_UMMA_DEBUG bits: 1 no MMA, 2 no TMA, 4 no epilogue) to see which role bounds it.
Each configuration runs in a subprocess (the flag is read once per process)."""
import os, subprocess, sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

SHAPES = [(8, 96, 96, 256, 256), (8, 192, 192, 128, 128), (8, 384, 384, 64, 64)]

def one(flag, mode):
    import torch
    from sda_b200 import _lib
    lib = _lib.load()
    for (N, Cin, Cout, H, W) in SHAPES:
        x = torch.randn(N, Cin, H, W, device='cuda'); w = torch.randn(Cout, Cin, 3, 3, device='cuda') * 0.03
        b = torch.randn(Cout, device='cuda')
        nbytes = lib.sdab_conv3x3_workspace_bytes(N, Cin, Cout, H, W, 1, 0)
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device='cuda'); base = (ws.data_ptr() + 1023) // 1024 * 1024
        out = torch.empty(N, Cout, H, W, device='cuda')
        def run():
            _lib.check(lib.sdab_conv3x3(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), N, Cin, Cout, H, W, 1, 0, mode, 0, base, nbytes, _lib.stream_ptr()))
        for _ in range(3): run()
        torch.cuda.synchronize()
        lib.sdab_conv_profile(1)
        for _ in range(5): run()
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        _lib.check(lib.sdab_conv_profile_read(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)))
        lib.sdab_conv_profile(0)
        t = ms.value / n.value
        tiles = N * H * W / 128
        stages = 9 * Cin / 32
        cyc = t * 1e-3 * 1.9e9 / (tiles / 148) / stages
        print(f'flag={flag} mode={mode} {Cin}->{Cout} {H}x{W}: {t:.3f} ms  {fl.value / n.value / t / 1e9:.0f} TF/s alg  ~{cyc:.0f} cyc/stage', flush=True)

if __name__ == '__main__':
    if len(sys.argv) > 1:
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for mode in (0, 1):
            for flag in (0, 4, 2):
                env = dict(os.environ, SDAB_UMMA_DEBUG=str(flag))
                r = subprocess.run([sys.executable, __file__, str(flag), str(mode)], env=env, capture_output=True, text=True, timeout=120)
                print(r.stdout, r.stderr[-500:] if r.returncode else '', flush=True)
