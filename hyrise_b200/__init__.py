"""hyrise_b200 — Blackwell-native (sm_100a) execution path for Hyrise's TableScan / JoinHash / AggregateHash.

The compute lives in ``hyrise_b200/lib/libhyrise_b200.so`` (hand-written CUDA behind the C-ABI of
``include/hyrise_b200.h``); this package is the host-side mirror of the reference's storage and operator interface used
by the tests and by bench.py. There is no CPU fallback: without the built library the import of ``capi.load_library``
users fails.
"""
from . import capi  # noqa: F401
from .storage import (ColumnDefinition, Chunk, Segment, Table, load_table, encode_dictionary,  # noqa: F401
                      encode_frame_of_reference, make_value_segment)

__all__ = ["capi", "ColumnDefinition", "Chunk", "Segment", "Table", "load_table"]
