r"""Drop-in alias: `import sda` resolves to the B200-native package `sda_b200`, so the
reference's experiments/ scripts (`from sda.mcs import *`, `from sda.score import *`,
`from sda.utils import *`) run unchanged."""

import sys

import sda_b200
from sda_b200 import mcs, nn, score, utils  # noqa: F401

for _name in ('mcs', 'nn', 'score', 'utils'):
    sys.modules[f'{__name__}.{_name}'] = getattr(sda_b200, _name)

__version__ = sda_b200.__version__
