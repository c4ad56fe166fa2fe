r"""sda_b200 -- B200-native implementation of the SDA hot path.

Drop-in for the reference package `sda` (francois-rozet/sda): same submodules, class
names and call signatures for the batched Markov-blanket score evaluation
(`score`, `nn`) and the Kolmogorov-flow stepper (`mcs`), executed by hand-written
sm_100a CUDA kernels behind the C ABI of `include/sdab.h` (`libsdab.so`).
"""

from . import mcs, nn, score, utils  # noqa: F401

__version__ = '0.1.0'
