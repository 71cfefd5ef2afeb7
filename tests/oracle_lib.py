"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY — never imported by hyrise_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from hyrise_b200 import capi
from hyrise_b200.device import (ROW_ID_DTYPE, Aggregate, AggregateOutput, Predicate, build_aggregate_defs,
                                build_scan_predicate)
from hyrise_b200.storage import NUMPY_TYPES, Table

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(REPO, "oracle", "liboracle.so")


class OrcPosList(C.Structure):
    _fields_ = [("chunk_count", C.c_uint32), ("total", C.c_uint64), ("chunk_offsets", C.POINTER(C.c_uint64)),
                ("row_ids", C.POINTER(capi.RowID))]


class OrcJoinResult(C.Structure):
    _fields_ = [
        ("pair_count", C.c_uint64),
        ("build_row_ids", C.POINTER(capi.RowID)),
        ("probe_row_ids", C.POINTER(capi.RowID)),
        ("radix_bits", C.c_int32),
        ("partition_count", C.c_uint32),
        ("partition_offsets", C.POINTER(C.c_uint64)),
        ("slice_count", C.c_uint32),
        ("slice_offsets", C.POINTER(C.c_uint64)),
        ("output_chunk_count", C.c_uint32),
        ("output_chunk_offsets", C.POINTER(C.c_uint64)),
        ("build_materialized", C.c_uint64),
        ("probe_materialized", C.c_uint64),
    ]


class OrcAggregateColumn(C.Structure):
    _fields_ = [("value_type", C.c_int32), ("values", C.c_void_p), ("nulls", C.POINTER(C.c_uint8))]


class OrcAggregateResult(C.Structure):
    _fields_ = [("group_count", C.c_uint64), ("used_immediate_keys", C.c_int32), ("row_ids", C.POINTER(capi.RowID)),
                ("aggregate_count", C.c_uint32), ("columns", C.POINTER(OrcAggregateColumn))]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_PATH):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")])
    lib = C.CDLL(ORACLE_PATH)
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_calculate_radix_bits.restype = C.c_int32
    lib.orc_calculate_radix_bits.argtypes = [C.c_uint64, C.c_uint64]
    lib.orc_table_scan.argtypes = [C.POINTER(capi.TableView), C.POINTER(capi.ScanPredicate), C.POINTER(OrcPosList),
                                   C.c_int32, C.POINTER(OrcPosList)]
    lib.orc_pos_list_free.argtypes = [C.POINTER(OrcPosList)]
    lib.orc_pos_list_free.restype = None
    lib.orc_join_hash.argtypes = [C.POINTER(capi.TableView), C.c_uint32, C.POINTER(OrcPosList),
                                  C.POINTER(capi.TableView), C.c_uint32, C.POINTER(OrcPosList), C.c_int32, C.c_int32,
                                  C.c_int32, C.POINTER(OrcJoinResult)]
    lib.orc_join_result_free.argtypes = [C.POINTER(OrcJoinResult)]
    lib.orc_join_result_free.restype = None
    lib.orc_aggregate_hash.argtypes = [C.POINTER(capi.TableView), C.POINTER(capi.AggregateQuery), C.POINTER(OrcPosList),
                                       C.c_int32, C.c_int32, C.POINTER(OrcAggregateResult)]
    lib.orc_aggregate_result_free.argtypes = [C.POINTER(OrcAggregateResult)]
    lib.orc_aggregate_result_free.restype = None
    lib.orc_encode_dictionary.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                          C.POINTER(C.c_uint32), C.c_void_p]
    lib.orc_compress_fixed_width.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_int32)]
    lib.orc_compress_bitpacking.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_int32)]
    lib.orc_encode_frame_of_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                                  C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    lib.orc_decode_segment.argtypes = [C.POINTER(capi.SegmentDesc), C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


class OracleError(RuntimeError):
    pass


def _check(status: int) -> None:
    if status != 0:
        raise OracleError(f"status {status}: {load().orc_last_error().decode()}")


class HostPosList:
    """Flat RowID list + per-chunk offsets (what hyb_pos_list_* / orc_pos_list hold)."""

    def __init__(self, row_ids: np.ndarray, chunk_offsets: np.ndarray):
        self.row_ids = np.ascontiguousarray(row_ids, dtype=ROW_ID_DTYPE)
        self.chunk_offsets = np.ascontiguousarray(chunk_offsets, dtype=np.uint64)

    def struct(self) -> OrcPosList:
        s = OrcPosList()
        s.chunk_count = len(self.chunk_offsets) - 1
        s.total = len(self.row_ids)
        s.chunk_offsets = self.chunk_offsets.ctypes.data_as(C.POINTER(C.c_uint64))
        s.row_ids = C.cast(self.row_ids.ctypes.data, C.POINTER(capi.RowID))
        return s

    def chunk(self, chunk_id: int) -> np.ndarray:
        return self.row_ids[int(self.chunk_offsets[chunk_id]): int(self.chunk_offsets[chunk_id + 1])]


def _copy_row_ids(pointer, count: int) -> np.ndarray:
    if count == 0 or not pointer:
        return np.zeros(0, dtype=ROW_ID_DTYPE)
    return np.ctypeslib.as_array(C.cast(pointer, C.POINTER(C.c_uint64)), shape=(count,)).view(ROW_ID_DTYPE).copy()


def _copy_u64(pointer, count: int) -> np.ndarray:
    return np.ctypeslib.as_array(pointer, shape=(count,)).copy()


def table_scan(table: Table, predicate: Predicate, input_filter: HostPosList | None = None, threads: int = 1,
               ) -> HostPosList:
    lib = load()
    holder = table.view()
    struct, keepalive = build_scan_predicate(table, predicate)
    filter_struct = input_filter.struct() if input_filter is not None else None
    out = OrcPosList()
    _check(lib.orc_table_scan(holder.pointer(), C.byref(struct), C.byref(filter_struct) if filter_struct else None,
                              threads, C.byref(out)))
    result = HostPosList(_copy_row_ids(out.row_ids, out.total), _copy_u64(out.chunk_offsets, out.chunk_count + 1))
    lib.orc_pos_list_free(C.byref(out))
    del keepalive
    return result


class HostJoinResult:
    def __init__(self, raw: OrcJoinResult):
        n = raw.pair_count
        self.pair_count = n
        self.build = _copy_row_ids(raw.build_row_ids, n) if raw.build_row_ids else None
        self.probe = _copy_row_ids(raw.probe_row_ids, n)
        self.radix_bits = raw.radix_bits
        self.partition_offsets = _copy_u64(raw.partition_offsets, raw.partition_count + 1)
        self.slice_offsets = _copy_u64(raw.slice_offsets, raw.slice_count + 1)
        self.output_chunk_offsets = _copy_u64(raw.output_chunk_offsets, raw.output_chunk_count + 1)
        self.build_materialized = raw.build_materialized
        self.probe_materialized = raw.probe_materialized


def join_hash(build: Table, build_column: int, probe: Table, probe_column: int, mode: int, radix_bits: int = -1,
              threads: int = 1, build_filter: HostPosList | None = None, probe_filter: HostPosList | None = None,
              ) -> HostJoinResult:
    lib = load()
    build_holder, probe_holder = build.view(), probe.view()
    build_struct = build_filter.struct() if build_filter is not None else None
    probe_struct = probe_filter.struct() if probe_filter is not None else None
    raw = OrcJoinResult()
    _check(lib.orc_join_hash(build_holder.pointer(), build_column, C.byref(build_struct) if build_struct else None,
                             probe_holder.pointer(), probe_column, C.byref(probe_struct) if probe_struct else None,
                             mode, radix_bits, threads, C.byref(raw)))
    result = HostJoinResult(raw)
    lib.orc_join_result_free(C.byref(raw))
    return result


def aggregate_hash(table: Table, groupby_column_ids, aggregates: list[Aggregate], predicates=(),
                   input_filter: HostPosList | None = None, threads: int = 1, parallel: bool = False) -> AggregateOutput:
    lib = load()
    holder = table.view()
    query = capi.AggregateQuery()
    keepalive = []
    predicate_structs = (capi.ScanPredicate * max(len(predicates), 1))()
    for index, predicate in enumerate(predicates):
        struct, alive = build_scan_predicate(table, predicate)
        predicate_structs[index] = struct
        keepalive.append(alive)
    query.predicate_count = len(predicates)
    query.predicates = C.cast(predicate_structs, C.POINTER(capi.ScanPredicate))
    groupby = (C.c_uint32 * max(len(groupby_column_ids), 1))(*groupby_column_ids)
    query.groupby_count = len(groupby_column_ids)
    query.groupby_column_ids = C.cast(groupby, C.POINTER(C.c_uint32))
    defs = build_aggregate_defs(aggregates)
    query.aggregate_count = len(aggregates)
    query.aggregates = C.cast(defs, C.POINTER(capi.AggregateDef))
    filter_struct = input_filter.struct() if input_filter is not None else None
    raw = OrcAggregateResult()
    _check(lib.orc_aggregate_hash(holder.pointer(), C.byref(query), C.byref(filter_struct) if filter_struct else None,
                                  threads, 1 if parallel else 0, C.byref(raw)))
    count = raw.group_count
    row_ids = _copy_row_ids(raw.row_ids, count)
    values, nulls, types = [], [], []
    for index in range(raw.aggregate_count):
        column = raw.columns[index]
        dtype = np.dtype(NUMPY_TYPES[column.value_type])
        if count:
            buffer = (C.c_uint8 * (count * dtype.itemsize)).from_address(column.values)
            values.append(np.frombuffer(buffer, dtype=dtype, count=count).copy())
            nulls.append(np.ctypeslib.as_array(column.nulls, shape=(count,)).astype(bool))
        else:
            values.append(np.zeros(0, dtype=dtype))
            nulls.append(np.zeros(0, dtype=bool))
        types.append(column.value_type)
    output = AggregateOutput(count, bool(raw.used_immediate_keys), row_ids, values, nulls, types)
    lib.orc_aggregate_result_free(C.byref(raw))
    return output


# --- encoders (used to cross-check hyrise_b200.storage) -----------------------------------------------------------
def encode_dictionary(values: np.ndarray, nulls: np.ndarray | None, data_type: int):
    lib = load()
    values = np.ascontiguousarray(values, dtype=NUMPY_TYPES[data_type])
    n = len(values)
    null_bytes = None if nulls is None else np.ascontiguousarray(nulls, dtype=np.uint8)
    dictionary = np.empty(max(n, 1), dtype=values.dtype)
    ids = np.empty(max(n, 1), dtype=np.uint32)
    size = C.c_uint32()
    _check(lib.orc_encode_dictionary(data_type, values.ctypes.data, None if null_bytes is None else null_bytes.ctypes.data,
                                     n, dictionary.ctypes.data, C.byref(size), ids.ctypes.data))
    return dictionary[: size.value].copy(), ids[:n].copy()


def compress_fixed_width(ids: np.ndarray, max_value: int):
    lib = load()
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    out = np.empty(max(len(ids), 1) * 4, dtype=np.uint8)
    vector_type = C.c_int32()
    _check(lib.orc_compress_fixed_width(ids.ctypes.data, len(ids), max_value, out.ctypes.data, C.byref(vector_type)))
    dtype = {capi.VEC_FIXED_1B: np.uint8, capi.VEC_FIXED_2B: np.uint16, capi.VEC_FIXED_4B: np.uint32}[vector_type.value]
    return out.view(dtype)[: len(ids)].copy(), vector_type.value


def compress_bitpacking(ids: np.ndarray):
    lib = load()
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    out = np.zeros((len(ids) * 32 + 63) // 64 + 1, dtype=np.uint64)
    bits = C.c_int32()
    _check(lib.orc_compress_bitpacking(ids.ctypes.data, len(ids), out.ctypes.data, C.byref(bits)))
    return out[: (len(ids) * bits.value + 63) // 64].copy(), bits.value


def encode_frame_of_reference(values: np.ndarray, nulls: np.ndarray | None):
    lib = load()
    values = np.ascontiguousarray(values, dtype=np.int32)
    n = len(values)
    null_bytes = None if nulls is None else np.ascontiguousarray(nulls, dtype=np.uint8)
    minima = np.empty(max((n + capi.FOR_BLOCK_SIZE - 1) // capi.FOR_BLOCK_SIZE, 1), dtype=np.int32)
    offsets = np.empty(max(n, 1), dtype=np.uint32)
    max_offset, has_nulls = C.c_uint32(), C.c_int32()
    _check(lib.orc_encode_frame_of_reference(values.ctypes.data, None if null_bytes is None else null_bytes.ctypes.data,
                                             n, minima.ctypes.data, offsets.ctypes.data, C.byref(max_offset),
                                             C.byref(has_nulls)))
    blocks = (n + capi.FOR_BLOCK_SIZE - 1) // capi.FOR_BLOCK_SIZE
    return minima[:blocks].copy(), offsets[:n].copy(), max_offset.value, bool(has_nulls.value)


def decode_segment(segment) -> tuple[np.ndarray, np.ndarray]:
    lib = load()
    desc = segment.desc()
    dtype = np.uint32 if segment.data_type == capi.TYPE_STRING else NUMPY_TYPES[segment.data_type]
    values = np.zeros(max(segment.row_count, 1), dtype=dtype)
    nulls = np.zeros(max(segment.row_count, 1), dtype=np.uint8)
    _check(lib.orc_decode_segment(C.byref(desc), values.ctypes.data, nulls.ctypes.data))
    return values[: segment.row_count], nulls[: segment.row_count].astype(bool)


def debug_materialize(table: Table, column: int, keep_nulls: bool, radix_bits: int, input_bloom_slots=None):
    """materialize_input<int,int> internals for the KATs of join_hash_steps_test.cpp:169-263."""
    lib = load()
    holder = table.view()
    rows = table.row_count
    values = np.zeros(max(rows, 1), dtype=np.int32)
    row_ids = np.zeros(max(rows, 1), dtype=ROW_ID_DTYPE)
    nulls = np.zeros(max(rows, 1), dtype=np.uint8)
    histograms = np.zeros(max(table.chunk_count, 1) * (1 << radix_bits), dtype=np.uint64)
    bloom = np.zeros(4096, dtype=np.uint32)
    count, bloom_count = C.c_uint64(), C.c_uint32()
    slots = None if input_bloom_slots is None else np.ascontiguousarray(input_bloom_slots, dtype=np.uint32)
    lib.orc_debug_materialize.argtypes = [C.POINTER(capi.TableView), C.c_uint32, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
    _check(lib.orc_debug_materialize(holder.pointer(), column, int(keep_nulls), radix_bits,
                                     None if slots is None else slots.ctypes.data, 0 if slots is None else len(slots),
                                     values.ctypes.data, row_ids.ctypes.data, nulls.ctypes.data, C.byref(count),
                                     histograms.ctypes.data, bloom.ctypes.data, C.byref(bloom_count)))
    n = count.value
    return (values[:n], row_ids[:n], nulls[:n].astype(bool),
            histograms.reshape(max(table.chunk_count, 1), 1 << radix_bits)[: table.chunk_count],
            bloom[: bloom_count.value])
