"""tehmm_b200 -- B200-native replacement for teHmm's multitrack-HMM hot path.

Module names mirror the reference package (teHmm): `_hmm` and `_emission` hold
the drop-in kernels with the reference's Cython signatures, `emission` and
`hmm` the model classes.  Everything numeric runs in libtehmm_b200.so (CUDA,
sm_100a); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import common  # noqa: F401


def build():
    from .build import build as _b
    return _b()
