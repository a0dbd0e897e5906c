"""sequila-native_b200 — B200-native interval-overlap join (`SET sequila.interval_join_algorithm TO cuda`).

Only the hot path of sequila-native is here: build-side index, probe, emit and
column gather as hand-written sm_100a CUDA kernels behind the C ABI in ``include/sequila_cuda.h``
(``libsequila_cuda.so``), plus the host-side mirror of the reference's operator surface.
There is no CPU fallback: importing :mod:`cuda_join` objects without the built library raises.
"""
from ._native import SequilaCudaError, LIB_PATH  # noqa: F401
from .cuda_join import CudaContext, CudaIndex, CudaStream  # noqa: F401
from .scan import CudaScan  # noqa: F401
from . import synth  # noqa: F401
from .session import Algorithm, SequilaConfig, apply_set, ParseAlgorithmError  # noqa: F401
from . import intervals  # noqa: F401
from .interval_join import IntervalJoinExec, HashJoinDesc, optimize  # noqa: F401

__all__ = ["CudaContext", "CudaIndex", "CudaStream", "CudaScan", "SequilaCudaError", "synth", "LIB_PATH", "Algorithm",
           "SequilaConfig", "apply_set", "ParseAlgorithmError", "intervals", "IntervalJoinExec", "HashJoinDesc",
           "optimize"]
