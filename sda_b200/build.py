r"""Builds libsdab.so (the sm_100a CUDA library behind the C ABI of include/sdab.h) in-tree.

    python -m sda_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels
to the GPU box with the repository snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / 'csrc'
LIB = HERE / 'libsdab.so'
OBJ = HERE / 'build'

SOURCES = ['api.cu', 'elementwise.cu', 'conv_simt.cu', 'conv_umma.cu', 'unet.cu', 'conv_api.cu', 'score_ops.cu', 'kolmogorov.cu', 'wgrad.cu', 'wgrad_umma.cu', 'peer.cu']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr', '--expt-extended-lambda',
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError('nvcc not found: libsdab cannot be built (there is no CPU fallback)')


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob('*')) + [HERE.parent / 'include' / 'sdab.h', Path(__file__)]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = OBJ / 'stamp'
    digest = _digest()

    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB

    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)

    # one builder at a time (the ranks of a torchrun job all call load()); the others wait, then
    # find the stamp up to date
    import fcntl

    lock = open(OBJ / 'lock', 'w')
    fcntl.flock(lock, fcntl.LOCK_EX)

    try:
        if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
            return LIB

        return _build_locked(nvcc, stamp, digest, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc: str, stamp: Path, digest: str, verbose: bool) -> Path:

    def compile_one(src: str) -> Path:
        obj = OBJ / (src + '.o')
        cmd = [nvcc, *NVCC_FLAGS, '-c', str(CSRC / src), '-o', str(obj)]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(compile_one, SOURCES))

    cmd = [nvcc, '-shared', '-o', str(LIB), *map(str, objs), '-cudart', 'static', '-Xcompiler', '-fPIC']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')

    stamp.write_text(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
