"""coslam_b200 -- B200-native (sm_100a) implementation of CoSLAM's two data-parallel hot paths:
the pyramidal KLT tracker and multi-camera bundle adjustment / pose refinement.

The compute lives in csrc/ (hand-written CUDA behind the C-ABI of include/coslam_b200.h, built
into coslam_b200/libcoslam_b200.so).  This package is only the thin host-side binding used by the
tests and bench.py; it has no CPU fallback: importing `coslam_b200.api` fails loudly when the
shared library is missing."""
__version__ = "0.1"
