"""Shared helpers for the parity tests."""
from __future__ import annotations

import os

import numpy as np

from hyrise_b200 import capi
from hyrise_b200.device import ROW_ID_DTYPE, Predicate
from hyrise_b200.storage import ColumnDefinition, Table, load_table

HERE = os.path.dirname(os.path.abspath(__file__))
TBL = os.path.join(HERE, "golden", "tbl")
TPCH = os.path.join(HERE, "golden", "tpch")

ENCODINGS = ["Unencoded", "Dictionary", "FrameOfReference"]  # table_scan_test.cpp:267-270 minus RunLength


def tbl(name: str, chunk_size: int = capi.DEFAULT_CHUNK_SIZE) -> Table:
    return load_table(os.path.join(TBL, name), chunk_size)


def column_values_at(table: Table, column_id: int, row_ids: np.ndarray):
    """Dereference RowIDs like a ReferenceSegment (reference_segment.hpp): list of python values, None for NULL."""
    out = []
    cache = {}
    for row in row_ids:
        chunk_id, offset = int(row["chunk_id"]), int(row["chunk_offset"])
        if chunk_id not in cache:
            segment = table.chunks[chunk_id].segments[column_id]
            cache[chunk_id] = (segment.decode(), segment.null_mask())
        values, nulls = cache[chunk_id]
        out.append(None if nulls[offset] else values[offset].item())
    return out


def row_ids_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a, dtype=ROW_ID_DTYPE)
    b = np.ascontiguousarray(b, dtype=ROW_ID_DTYPE)
    return a.shape == b.shape and bool(np.array_equal(a.view(np.uint64), b.view(np.uint64)))


def assert_pos_lists_equal(device_rows, device_offsets, oracle_list) -> None:
    assert np.array_equal(np.asarray(device_offsets, dtype=np.uint64), oracle_list.chunk_offsets), \
        "per-chunk PosList boundaries differ"
    assert row_ids_equal(device_rows, oracle_list.row_ids), "RowIDs differ"


def random_table(rng: np.random.Generator, rows: int, chunk_size: int, with_nulls: bool = True) -> Table:
    """Five-type table (int, long, float, double + a low-cardinality int) for randomized parity runs."""
    definitions = [
        ColumnDefinition("i", capi.TYPE_INT32, with_nulls),
        ColumnDefinition("l", capi.TYPE_INT64, with_nulls),
        ColumnDefinition("f", capi.TYPE_FLOAT32, with_nulls),
        ColumnDefinition("d", capi.TYPE_FLOAT64, with_nulls),
        ColumnDefinition("k", capi.TYPE_INT32, False),
    ]
    columns = [
        rng.integers(-1000, 1000, rows, dtype=np.int32),
        rng.integers(-(10 ** 12), 10 ** 12, rows, dtype=np.int64),
        (rng.integers(0, 2000, rows) / 8).astype(np.float32),
        rng.normal(0, 100, rows).astype(np.float64),
        rng.integers(0, 300, rows, dtype=np.int32),
    ]
    nulls = [(rng.random(rows) < 0.1) if (with_nulls and d.nullable) else None for d in definitions]
    return Table.from_columns(definitions, columns, nulls, chunk_size)
