"""Thin object wrapper over the C-ABI (capi.py): context, device tables, device-resident results.

Everything here is plumbing: it marshals numpy buffers into the structs of include/hyrise_b200.h and back. All compute
happens in libhyrise_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import capi
from .capi import check
from .storage import NUMPY_TYPES, Table, TableViewHolder

ROW_ID_DTYPE = np.dtype([("chunk_id", np.uint32), ("chunk_offset", np.uint32)])


@dataclass
class Predicate:
    """column <condition> value | column BETWEEN lower AND upper | column IS [NOT] NULL — the arguments of
    ColumnVsValueTableScanImpl / ColumnBetweenTableScanImpl / ColumnIsNullTableScanImpl after
    lossless_predicate_variant_cast (table_scan.cpp:340-366)."""

    column_id: int
    condition: int
    lower: object = None
    upper: object = None


def _set_value(target: capi.Value, data_type: int, value) -> None:
    if value is None:
        return
    if data_type == capi.TYPE_INT32:
        target.i32 = int(value)
    elif data_type == capi.TYPE_INT64:
        target.i64 = int(value)
    elif data_type == capi.TYPE_FLOAT32:
        target.f32 = float(np.float32(value))
    elif data_type == capi.TYPE_FLOAT64:
        target.f64 = float(value)


def build_scan_predicate(table: Table, predicate: Predicate) -> tuple[capi.ScanPredicate, object]:
    """Returns the C struct plus an object that must stay alive while the struct is in use."""
    data_type = table.column_definitions[predicate.column_id].data_type
    struct = capi.ScanPredicate()
    struct.column_id = predicate.column_id
    struct.condition = predicate.condition
    keepalive = None
    needs_value = predicate.condition not in (capi.PRED_IS_NULL, capi.PRED_IS_NOT_NULL)
    if data_type == capi.TYPE_STRING:
        if needs_value:
            keepalive = np.ascontiguousarray(table.string_value_id_bounds(predicate), dtype=np.uint32)
            struct.value_id_bounds = keepalive.ctypes.data
    else:
        _set_value(struct.lower, data_type, predicate.lower)
        _set_value(struct.upper, data_type, predicate.upper)
    return struct, keepalive


@dataclass
class Expression:
    """Reverse-Polish arithmetic over columns/literals: nodes are ("col", id) | ("lit", type, value) | "+" "-" "*" "/"."""

    nodes: Sequence

    @staticmethod
    def column(column_id: int) -> "Expression":
        return Expression([("col", column_id)])


@dataclass
class Aggregate:
    function: int
    expression: Expression | None = None  # None for COUNT(*)


_OPS = {"+": capi.EXPR_ADD, "-": capi.EXPR_SUB, "*": capi.EXPR_MUL, "/": capi.EXPR_DIV}


def build_aggregate_defs(aggregates: Sequence[Aggregate]):
    defs = (capi.AggregateDef * max(len(aggregates), 1))()
    for index, aggregate in enumerate(aggregates):
        defs[index].function = aggregate.function
        nodes = [] if aggregate.expression is None else list(aggregate.expression.nodes)
        if len(nodes) > capi.MAX_EXPR_NODES:
            raise capi.UnsupportedOnDevice(capi.HYB_ERR_UNSUPPORTED, "expression too long")
        defs[index].node_count = len(nodes)
        for position, node in enumerate(nodes):
            target = defs[index].nodes[position]
            if isinstance(node, str):
                target.op = _OPS[node]
            elif node[0] == "col":
                target.op = capi.EXPR_COLUMN
                target.column_id = node[1]
            else:
                target.op = capi.EXPR_LITERAL
                target.literal_type = node[1]
                _set_value(target.literal, node[1], node[2])
    return defs


class DeviceTable:
    def __init__(self, context: "DeviceContext", handle: int, host_table: Table, holder: TableViewHolder):
        self.context = context
        self.handle = handle
        self.host_table = host_table
        self._holder = holder

    def drop(self) -> None:
        if self.handle:
            check(self.context.lib.hyb_table_drop(self.context.ptr, self.handle))
            self.handle = 0


class DevicePosList:
    """Device-resident output of hyb_table_scan: a RowIDPosList per input chunk."""

    def __init__(self, context: "DeviceContext", handle: int):
        self.context = context
        self.handle = handle

    def info(self) -> tuple[int, int]:
        total, chunks = C.c_uint64(), C.c_uint32()
        check(self.context.lib.hyb_pos_list_info(self.context.ptr, self.handle, C.byref(total), C.byref(chunks)))
        return total.value, chunks.value

    def chunk_offsets(self) -> np.ndarray:
        _, chunks = self.info()
        offsets = np.empty(chunks + 1, dtype=np.uint64)
        check(self.context.lib.hyb_pos_list_chunk_offsets(self.context.ptr, self.handle, offsets.ctypes.data))
        return offsets

    def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        total, _ = self.info()
        if out is None:
            out = np.empty(total, dtype=ROW_ID_DTYPE)
        check(self.context.lib.hyb_pos_list_copy(self.context.ptr, self.handle, 0, total, out.ctypes.data))
        return out[:total]

    def free(self) -> None:
        if self.handle:
            check(self.context.lib.hyb_pos_list_free(self.context.ptr, self.handle))
            self.handle = 0


class DeviceJoinResult:
    def __init__(self, context: "DeviceContext", handle: int, has_build_side: bool):
        self.context = context
        self.handle = handle
        self.has_build_side = has_build_side

    def info(self) -> tuple[int, int, int]:
        pairs, partitions, bits = C.c_uint64(), C.c_uint32(), C.c_int32()
        check(self.context.lib.hyb_join_result_info(self.context.ptr, self.handle, C.byref(pairs), C.byref(partitions),
                                                     C.byref(bits)))
        return pairs.value, partitions.value, bits.value

    def partition_offsets(self) -> np.ndarray:
        _, partitions, _ = self.info()
        offsets = np.empty(partitions + 1, dtype=np.uint64)
        check(self.context.lib.hyb_join_result_partition_offsets(self.context.ptr, self.handle, offsets.ctypes.data))
        return offsets

    def to_host(self, out_build: np.ndarray | None = None, out_probe: np.ndarray | None = None):
        pairs, _, _ = self.info()
        probe = np.empty(pairs, dtype=ROW_ID_DTYPE) if out_probe is None else out_probe
        build = None
        if self.has_build_side:
            build = np.empty(pairs, dtype=ROW_ID_DTYPE) if out_build is None else out_build
        check(self.context.lib.hyb_join_result_copy(self.context.ptr, self.handle, 0, pairs,
                                                    None if build is None else build.ctypes.data, probe.ctypes.data))
        return (None if build is None else build[:pairs]), probe[:pairs]

    def pos_list(self, side: int) -> "DevicePosList":
        """hyb_join_result_pos_list: one side (0 = build, 1 = probe) as a PosList on that side's table, in result order."""
        handle = C.c_uint64()
        check(self.context.lib.hyb_join_result_pos_list(self.context.ptr, self.handle, side, C.byref(handle)))
        return DevicePosList(self.context, handle.value)

    def output_chunks(self) -> np.ndarray:
        """hyb_join_result_output_chunks: offsets of the chunks write_output_chunks cuts the result into."""
        count = C.c_uint32()
        check(self.context.lib.hyb_join_result_output_chunks(self.context.ptr, self.handle, None, C.byref(count)))
        offsets = np.zeros(count.value + 1, dtype=np.uint64)
        check(self.context.lib.hyb_join_result_output_chunks(self.context.ptr, self.handle, offsets.ctypes.data, C.byref(count)))
        return offsets

    def free(self) -> None:
        if self.handle:
            check(self.context.lib.hyb_join_result_free(self.context.ptr, self.handle))
            self.handle = 0


@dataclass
class AggregateOutput:
    group_count: int
    used_immediate_keys: bool
    row_ids: np.ndarray
    values: list[np.ndarray]
    nulls: list[np.ndarray]
    value_types: list[int]
    result_handle: int = 0   # != 0: the device-side result is still alive (aggregate_hash(..., keep_result=True))


class DevicePeerGroup:
    """hyb_peer_group: this rank's exchange arena and its mappings of the peers' arenas. Distributed operators are one
    C-ABI call per rank; the host layer only moves the IPC handles once (hyrise_b200.distributed.connect_peer_group)."""

    def __init__(self, context: "DeviceContext", rank: int, world: int, tuple_capacity: int):
        self.context = context
        self.rank, self.world = rank, world
        handle = (C.c_ubyte * capi.IPC_HANDLE_BYTES)()
        group = C.c_uint64()
        check(context.lib.hyb_peer_group_create(context.ptr, rank, world, tuple_capacity, handle, C.byref(group)))
        self.handle = group.value
        self.ipc_handle = bytes(handle)

    def connect(self, all_ipc_handles: Sequence[bytes]) -> None:
        packed = (C.c_ubyte * (capi.IPC_HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(all_ipc_handles))
        check(self.context.lib.hyb_peer_group_connect(self.context.ptr, self.handle, packed))

    def join_hash(self, build: "DeviceTable", build_column: int, probe: "DeviceTable", probe_column: int,
                  build_chunk_base: int, probe_chunk_base: int, radix_bits: int = -1) -> "DeviceJoinResult":
        build_side = capi.JoinSide(build.handle, build_column, 0)
        probe_side = capi.JoinSide(probe.handle, probe_column, 0)
        handle = C.c_uint64()
        check(self.context.lib.hyb_join_hash_distributed(self.context.ptr, self.handle, C.byref(build_side), C.byref(probe_side),
                                                         build_chunk_base, probe_chunk_base, radix_bits, C.byref(handle)))
        return DeviceJoinResult(self.context, handle.value, True)

    def aggregate_hash(self, table: "DeviceTable", groupby_column_ids: Sequence[int], aggregates: Sequence["Aggregate"],
                       predicates: Sequence["Predicate"] = (), chunk_id_base: int = 0, position_base: int = 0) -> "AggregateOutput":
        query, keepalive = self.context._aggregate_query(table, groupby_column_ids, aggregates, predicates, None)
        handle = C.c_uint64()
        check(self.context.lib.hyb_aggregate_hash_distributed(self.context.ptr, self.handle, C.byref(query), chunk_id_base,
                                                              position_base, C.byref(handle)))
        del keepalive
        return self.context._collect_aggregate(handle.value, len(aggregates))

    def stats(self) -> capi.DistributedStats:
        stats = capi.DistributedStats()
        check(self.context.lib.hyb_peer_group_stats(self.context.ptr, self.handle, C.byref(stats)))
        return stats

    def destroy(self) -> None:
        if self.handle:
            check(self.context.lib.hyb_peer_group_destroy(self.context.ptr, self.handle))
            self.handle = 0


class DeviceContext:
    """hyb_context: one per process and GPU."""

    def __init__(self, device_index: int = 0):
        self.lib = capi.load_library()
        ptr = C.c_void_p()
        check(self.lib.hyb_context_create(device_index, C.byref(ptr)))
        self.ptr = ptr
        self.device_index = device_index

    def close(self) -> None:
        if self.ptr:
            check(self.lib.hyb_context_destroy(self.ptr))
            self.ptr = None

    def __enter__(self) -> "DeviceContext":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def synchronize(self) -> None:
        check(self.lib.hyb_context_synchronize(self.ptr))

    def set_option(self, name: str, value: str) -> None:
        """hyb_context_set_option: tuning / test knobs (the HYB_* environment variables are read at context creation)."""
        check(self.lib.hyb_context_set_option(self.ptr, name.encode(), str(value).encode()))

    # device column pool ------------------------------------------------------------------------------------------
    def upload(self, table) -> DeviceTable:
        """`table`: storage.Table or tpch.GeneratedTable (anything with view() / column_definitions)."""
        holder = table.view()
        handle = C.c_uint64()
        check(self.lib.hyb_table_upload(self.ptr, holder.pointer(), C.byref(handle)))
        return DeviceTable(self, handle.value, table, holder)

    def upload_blocks(self, blocks) -> int:
        """hyb_blocks_upload: DMA whole host arena blocks; returns the block-set handle."""
        array = (capi.HostBlock * max(len(blocks), 1))(*blocks)
        handle = C.c_uint64()
        check(self.lib.hyb_blocks_upload(self.ptr, array, len(blocks), C.byref(handle)))
        return handle.value

    def upload_from_blocks(self, table, block_set: int) -> DeviceTable:
        holder = table.view()
        handle = C.c_uint64()
        check(self.lib.hyb_table_upload_from_blocks(self.ptr, holder.pointer(), block_set, C.byref(handle)))
        return DeviceTable(self, handle.value, table, holder)

    def free_blocks(self, block_set: int) -> None:
        check(self.lib.hyb_blocks_free(self.ptr, block_set))

    # operators ---------------------------------------------------------------------------------------------------
    def table_scan(self, table: DeviceTable, predicate: Predicate, input_filter: DevicePosList | None = None,
                   ) -> DevicePosList:
        struct, keepalive = build_scan_predicate(table.host_table, predicate)
        handle = C.c_uint64()
        check(self.lib.hyb_table_scan(self.ptr, table.handle, C.byref(struct),
                                      input_filter.handle if input_filter else 0, C.byref(handle)))
        del keepalive
        return DevicePosList(self, handle.value)

    def join_hash(self, build: DeviceTable, build_column: int, probe: DeviceTable, probe_column: int, mode: int,
                  radix_bits: int = -1, build_filter: DevicePosList | None = None,
                  probe_filter: DevicePosList | None = None) -> DeviceJoinResult:
        build_side = capi.JoinSide(build.handle, build_column, build_filter.handle if build_filter else 0)
        probe_side = capi.JoinSide(probe.handle, probe_column, probe_filter.handle if probe_filter else 0)
        handle = C.c_uint64()
        check(self.lib.hyb_join_hash(self.ptr, C.byref(build_side), C.byref(probe_side), mode, radix_bits,
                                     C.byref(handle)))
        semi_or_anti = mode in (capi.JOIN_SEMI, capi.JOIN_ANTI_NULL_AS_TRUE, capi.JOIN_ANTI_NULL_AS_FALSE)
        return DeviceJoinResult(self, handle.value, not semi_or_anti)

    def _aggregate_query(self, table: DeviceTable, groupby_column_ids, aggregates, predicates, input_filter):
        """AggregateQuery struct + the objects that must stay alive while it is in use."""
        query = capi.AggregateQuery()
        query.table = table.handle
        query.filter = input_filter.handle if input_filter else 0
        keepalive = []
        predicate_structs = (capi.ScanPredicate * max(len(predicates), 1))()
        for index, predicate in enumerate(predicates):
            struct, alive = build_scan_predicate(table.host_table, predicate)
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
        keepalive += [predicate_structs, groupby, defs]
        return query, keepalive

    def aggregate_top_k(self, result_handle: int, aggregate_index: int, k: int, descending: bool = True) -> np.ndarray:
        """hyb_aggregate_result_top_k on a result kept alive with aggregate_hash(..., keep_result=True)."""
        indexes = np.zeros(k, dtype=np.uint32)
        count = C.c_uint32()
        check(self.lib.hyb_aggregate_result_top_k(self.ptr, result_handle, aggregate_index, k, int(descending),
                                                  indexes.ctypes.data, C.byref(count)))
        return indexes[: count.value]

    def free_aggregate_result(self, result_handle: int) -> None:
        check(self.lib.hyb_aggregate_result_free(self.ptr, result_handle))

    def _collect_aggregate(self, handle: int, aggregate_count: int, keep_result: bool = False) -> AggregateOutput:
        try:
            groups, immediate = C.c_uint64(), C.c_int32()
            check(self.lib.hyb_aggregate_result_info(self.ptr, handle, C.byref(groups), C.byref(immediate)))
            count = groups.value
            row_ids = np.empty(count, dtype=ROW_ID_DTYPE)
            check(self.lib.hyb_aggregate_result_row_ids(self.ptr, handle, row_ids.ctypes.data))
            values, nulls, types = [], [], []
            for index in range(aggregate_count):
                raw = np.zeros(max(count, 1), dtype=np.uint64)
                null = np.zeros(max(count, 1), dtype=np.uint8)
                value_type = C.c_int32()
                check(self.lib.hyb_aggregate_result_values(self.ptr, handle, index, raw.ctypes.data, null.ctypes.data,
                                                           C.byref(value_type)))
                dtype = NUMPY_TYPES[value_type.value]
                values.append(raw.view(np.uint8)[: count * np.dtype(dtype).itemsize].view(dtype).copy())
                nulls.append(null[:count].astype(bool))
                types.append(value_type.value)
        finally:
            if not keep_result:
                check(self.lib.hyb_aggregate_result_free(self.ptr, handle))
        output = AggregateOutput(count, bool(immediate.value), row_ids, values, nulls, types)
        output.result_handle = handle if keep_result else 0
        return output

    def aggregate_hash(self, table: DeviceTable, groupby_column_ids: Sequence[int], aggregates: Sequence[Aggregate],
                       predicates: Sequence[Predicate] = (), input_filter: DevicePosList | None = None,
                       keep_result: bool = False) -> AggregateOutput:
        query, keepalive = self._aggregate_query(table, groupby_column_ids, aggregates, predicates, input_filter)
        handle = C.c_uint64()
        check(self.lib.hyb_aggregate_hash(self.ptr, C.byref(query), C.byref(handle)))
        del keepalive
        return self._collect_aggregate(handle.value, len(aggregates), keep_result)

    def last_stats(self) -> capi.OperatorStats:
        stats = capi.OperatorStats()
        check(self.lib.hyb_last_operator_stats(self.ptr, C.byref(stats)))
        return stats

    # pinned host memory -------------------------------------------------------------------------------------------
    def pinned_empty(self, count: int, dtype) -> np.ndarray:
        """numpy array over hyb_host_alloc memory (freed with pinned_free)."""
        dtype = np.dtype(dtype)
        ptr = C.c_void_p()
        check(self.lib.hyb_host_alloc(max(count, 1) * dtype.itemsize, C.byref(ptr)))
        buffer = (C.c_uint8 * (max(count, 1) * dtype.itemsize)).from_address(ptr.value)
        array = np.frombuffer(buffer, dtype=dtype, count=count)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[array.ctypes.data] = ptr
        return array

    def pinned_free(self, array: np.ndarray) -> None:
        ptr = getattr(self, "_pinned", {}).pop(array.ctypes.data, None)
        if ptr is not None:
            check(self.lib.hyb_host_free(ptr))
