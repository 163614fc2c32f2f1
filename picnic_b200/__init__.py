"""picnic_b200 -- B200-native per-particle engine behind PICNIC's species/scattering
interfaces: CUDA kernels for sm_100a + a C ABI (include/picnic_gpu.h).

`capi` is the ctypes view of that ABI used by tests and bench.py; `decks` builds the
synthetic inputs.  The product has no CPU path: importing is cheap, but any call into
`capi` raises if libpicnic_gpu.so has not been built or no CUDA device is present.
"""
from . import decks  # noqa: F401

__all__ = ["decks", "capi", "build"]
