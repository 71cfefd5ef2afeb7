"""ctypes binding of the TPC-H-shaped workload generator (include/hyrise_b200_tpch.h) — bench/test tooling."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .storage import ColumnDefinition

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhyb_tpch.so")

L_ORDERKEY, L_QUANTITY, L_EXTENDEDPRICE, L_DISCOUNT, L_TAX, L_RETURNFLAG, L_LINESTATUS, L_SHIPDATE = range(8)
O_ORDERKEY, O_ORDERDATE = range(2)

LINEITEM_COLUMNS = [
    ColumnDefinition("l_orderkey", capi.TYPE_INT32), ColumnDefinition("l_quantity", capi.TYPE_FLOAT32),
    ColumnDefinition("l_extendedprice", capi.TYPE_FLOAT32), ColumnDefinition("l_discount", capi.TYPE_FLOAT32),
    ColumnDefinition("l_tax", capi.TYPE_FLOAT32), ColumnDefinition("l_returnflag", capi.TYPE_STRING),
    ColumnDefinition("l_linestatus", capi.TYPE_STRING), ColumnDefinition("l_shipdate", capi.TYPE_STRING),
]
ORDERS_COLUMNS = [ColumnDefinition("o_orderkey", capi.TYPE_INT32), ColumnDefinition("o_orderdate", capi.TYPE_STRING)]

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t)
FREE_FN = C.CFUNCTYPE(None, C.c_void_p)

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing; run `make` / __graft_entry__.build()")
        lib = C.CDLL(LIB_PATH)
        lib.hyb_tpch_generate.argtypes = [C.c_double, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_void_p)]
        lib.hyb_tpch_generate_shard.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p,
                                                C.POINTER(C.c_void_p)]
        lib.hyb_tpch_free.argtypes = [C.c_void_p]
        lib.hyb_tpch_free.restype = None
        for name in ("hyb_tpch_lineitem", "hyb_tpch_orders"):
            getattr(lib, name).argtypes = [C.c_void_p, C.POINTER(capi.TableView), C.POINTER(C.c_uint64)]
        lib.hyb_tpch_date_dictionary.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32,
                                                 C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_uint32)]
        lib.hyb_tpch_char_dictionary.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p),
                                                 C.POINTER(C.c_uint32)]
        lib.hyb_tpch_table_bytes.argtypes = [C.c_void_p, C.c_int32]
        lib.hyb_tpch_table_bytes.restype = C.c_uint64
        lib.hyb_tpch_value_id_bounds.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32,
                                                 C.c_void_p]
        lib.hyb_tpch_host_blocks.argtypes = [C.c_void_p, C.POINTER(capi.HostBlock), C.POINTER(C.c_uint32)]
        lib.hyb_tpch_day_number.argtypes = [C.c_int32] * 3
        lib.hyb_tpch_day_number.restype = C.c_int32
        _lib = lib
    return _lib


def day_number(date: str | bytes) -> int:
    """'YYYY-MM-DD' -> days since 1992-01-01."""
    text = date.decode() if isinstance(date, bytes) else date
    year, month, day = (int(part) for part in text.split("-"))
    return load().hyb_tpch_day_number(year, month, day)


class _ViewHolder:
    def __init__(self, view: capi.TableView):
        self.view = view

    def pointer(self):
        return C.byref(self.view)


class GeneratedTable:
    """A generated table: quacks like storage.Table for upload / predicates / the oracle binding."""

    def __init__(self, owner: "TpchTables", table_index: int, view: capi.TableView, rows: int, definitions):
        self.owner = owner
        self.table_index = table_index
        self._view = view
        self.row_count = rows
        self.column_definitions = definitions
        self.chunk_count = view.chunk_count
        self.column_count = view.column_count
        self._date_cache: dict[tuple[int, int], np.ndarray] = {}

    def view(self) -> _ViewHolder:
        return _ViewHolder(self._view)

    @property
    def host_bytes(self) -> int:
        return load().hyb_tpch_table_bytes(self.owner.ptr, self.table_index)

    def segment_desc(self, chunk_id: int, column_id: int) -> capi.SegmentDesc:
        return self._view.segments[chunk_id * self.column_count + column_id]

    def date_dictionary(self, column_id: int, chunk_id: int) -> np.ndarray:
        key = (column_id, chunk_id)
        if key not in self._date_cache:
            days, size = C.POINTER(C.c_int32)(), C.c_uint32()
            status = load().hyb_tpch_date_dictionary(self.owner.ptr, self.table_index, column_id, chunk_id,
                                                     C.byref(days), C.byref(size))
            if status != 0:
                raise ValueError("not a date column")
            self._date_cache[key] = np.ctypeslib.as_array(days, shape=(size.value,)) if size.value else \
                np.zeros(0, dtype=np.int32)
        return self._date_cache[key]

    def char_dictionary(self, column_id: int, chunk_id: int) -> bytes:
        chars, size = C.c_char_p(), C.c_uint32()
        status = load().hyb_tpch_char_dictionary(self.owner.ptr, column_id, chunk_id, C.byref(chars), C.byref(size))
        if status != 0:
            raise ValueError("not a char column")
        return chars.value[: size.value] if size.value else b""

    def value_id_at(self, column_id: int, chunk_id: int, offset: int) -> int:
        desc = self.segment_desc(chunk_id, column_id)
        width = {capi.VEC_FIXED_1B: C.c_uint8, capi.VEC_FIXED_2B: C.c_uint16, capi.VEC_FIXED_4B: C.c_uint32}[desc.vector_type]
        return C.cast(desc.attribute_vector, C.POINTER(width))[offset]

    def char_at(self, column_id: int, chunk_id: int, offset: int) -> int:
        """Byte value of a one-char string column at a row (group key of l_returnflag / l_linestatus)."""
        return self.char_dictionary(column_id, chunk_id)[self.value_id_at(column_id, chunk_id, offset)]

    def string_value_id_bounds(self, predicate) -> np.ndarray:
        """DictionarySegment::lower_bound / upper_bound per chunk (dictionary_segment.cpp:94-119) for date columns:
        day numbers order like the ISO strings."""
        between = capi.PRED_BETWEEN_INCLUSIVE <= predicate.condition <= capi.PRED_BETWEEN_EXCLUSIVE
        values = [predicate.lower, predicate.upper] if between else [predicate.lower]
        needles = (C.c_int32 * len(values))(*[day_number(value) if isinstance(value, (str, bytes)) else int(value)
                                             for value in values])
        bounds = np.empty((self.chunk_count, 2 * len(values)), dtype=np.uint32)
        status = load().hyb_tpch_value_id_bounds(self.owner.ptr, self.table_index, predicate.column_id, needles,
                                                 len(values), bounds.ctypes.data)
        if status != 0:
            raise ValueError("not a date column")
        return bounds


class TpchTables:
    def __init__(self, scale_factor: float, seed: int = 42, threads: int = 0, pinned: bool = False,
                 first_order: int = 0):
        lib = load()
        self._callbacks = None
        alloc = free = None
        if pinned:
            hyb = capi.load_library()

            def _alloc(size):
                ptr = C.c_void_p()
                return ptr.value if hyb.hyb_host_alloc(size, C.byref(ptr)) == 0 else None

            def _free(ptr):
                hyb.hyb_host_free(ptr)

            self._callbacks = (ALLOC_FN(_alloc), FREE_FN(_free))
            alloc, free = (C.cast(cb, C.c_void_p) for cb in self._callbacks)
        ptr = C.c_void_p()
        status = lib.hyb_tpch_generate_shard(scale_factor, seed, first_order, threads, alloc, free, C.byref(ptr))
        if status != 0:
            raise MemoryError(f"hyb_tpch_generate failed with status {status}")
        self.ptr = ptr
        self.scale_factor = scale_factor
        view, rows = capi.TableView(), C.c_uint64()
        lib.hyb_tpch_lineitem(ptr, C.byref(view), C.byref(rows))
        self.lineitem = GeneratedTable(self, 0, view, rows.value, LINEITEM_COLUMNS)
        view, rows = capi.TableView(), C.c_uint64()
        lib.hyb_tpch_orders(ptr, C.byref(view), C.byref(rows))
        self.orders = GeneratedTable(self, 1, view, rows.value, ORDERS_COLUMNS)

    def host_blocks(self) -> list:
        """The arena blocks behind all segment buffers (for DeviceContext.upload_blocks)."""
        count = C.c_uint32()
        load().hyb_tpch_host_blocks(self.ptr, None, C.byref(count))
        blocks = (capi.HostBlock * max(count.value, 1))()
        load().hyb_tpch_host_blocks(self.ptr, blocks, C.byref(count))
        return [capi.HostBlock(blocks[i].base, blocks[i].bytes) for i in range(count.value)]

    def close(self) -> None:
        if self.ptr:
            load().hyb_tpch_free(self.ptr)
            self.ptr = None
