"""dsk_b200 -- B200-native DSK counting hot path.

csrc/      hand-written CUDA (sm_100a) + the C ABI declared in include/dskgpu.h  -> libdskgpu.so
_lib.py    ctypes binding of the ABI (fails loudly if the library is missing; no CPU fallback)
counter.py object wrapper over one context
sorting_count.py  host-side mirror of gatb-core's SortingCountAlgorithm for this path
"""
from .counter import GpuCounter, DskGpuError  # noqa: F401
from .sorting_count import (SortingCountAlgorithm, BankStrings, BankFile, BankBytes, BankAlbum, open_bank,  # noqa: F401
                            getDefaultProperties)
