"""gangpack-b200: B200-native gang-scheduling bin-packer for the placement hot path of
palantir/k8s-spark-scheduler (tightly-pack / distribute-evenly + the FIFO fit-earlier-drivers loop).

Layout:
  csrc/          sm_100a CUDA kernels + the C ABI implementation (include/gangpack.h)
  native.py      ctypes binding of libgangpack.so (no CPU fallback)
  synth.py       seeded synthetic clusters / app queues
  host/          C++ host-side mirror of the reference's plug-in interface (binpacker.Binpacker ...)
  go/            the cgo shim a reference maintainer would add (source only; no Go toolchain here)
"""
from . import native, synth  # noqa: F401
from .native import (DISTRIBUTE_EVENLY, MINIMAL_FRAGMENTATION, MODE_FIFO_EXACT, MODE_FIFO_REFERENCE, MODE_INDEPENDENT,  # noqa: F401
                     TIGHTLY_PACK, GangPacker, GangpackError, MultiGangPacker)

__all__ = ["native", "synth", "GangPacker", "MultiGangPacker", "GangpackError", "TIGHTLY_PACK", "DISTRIBUTE_EVENLY", "MINIMAL_FRAGMENTATION",
           "MODE_INDEPENDENT", "MODE_FIFO_REFERENCE", "MODE_FIFO_EXACT"]
