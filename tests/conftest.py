import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault('SDAB_NO_TQDM', '1')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box')


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return

    skip = pytest.mark.skip(reason='no CUDA device')

    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return np.load(ROOT / 'tests' / 'golden' / f'{name}.npz')

    return load
