"""Pins the oracle's TableScan restatement to the reference's own expectations
(src/test/lib/operators/table_scan_test.cpp). CPU only."""
import numpy as np
import pytest

import oracle_lib as orc
from helpers import ENCODINGS, column_values_at, tbl
from hyrise_b200 import capi
from hyrise_b200.device import Predicate
from hyrise_b200.storage import ColumnDefinition, Table

P = capi


def scan_values(table, column_id, condition, value, out_column, upper=None, input_filter=None):
    result = orc.table_scan(table, Predicate(column_id, condition, value, upper), input_filter)
    return sorted(v for v in column_values_at(table, out_column, result.row_ids)), result


def int_int_tables(encoding):
    # table_scan_test.cpp:45-62: chunk sizes 7 / 5, first two chunks encoded
    compressed = tbl("int_int_shuffled.tbl", 7).encode(encoding, [0, 1])
    partly = tbl("int_int_shuffled_2.tbl", 5).encode(encoding, [0, 1])
    return compressed, partly


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_scan_on_compressed_segments(encoding):
    # table_scan_test.cpp:407-431
    expectations = {
        P.PRED_EQUALS: [106, 106],
        P.PRED_NOT_EQUALS: [100, 102, 104, 108, 110, 112] * 2,
        P.PRED_LESS_THAN: [100, 102, 104] * 2,
        P.PRED_LESS_THAN_EQUALS: [100, 102, 104, 106] * 2,
        P.PRED_GREATER_THAN: [108, 110, 112] * 2,
        P.PRED_GREATER_THAN_EQUALS: [106, 108, 110, 112] * 2,
        P.PRED_IS_NULL: [],
        P.PRED_IS_NOT_NULL: [100, 102, 104, 106, 108, 110, 112] * 2,
    }
    for table in int_int_tables(encoding):
        for condition, expected in expectations.items():
            values, _ = scan_values(table, 0, condition, 6, 1)
            assert values == sorted(expected), (encoding, condition)


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_scan_on_referenced_compressed_segments(encoding):
    # table_scan_test.cpp:433-463: scan(b < 108) then scan(a <op> 4) on the reference table
    expectations = {
        P.PRED_EQUALS: [104, 104],
        P.PRED_NOT_EQUALS: [100, 102, 106] * 2,
        P.PRED_LESS_THAN: [100, 102] * 2,
        P.PRED_LESS_THAN_EQUALS: [100, 102, 104] * 2,
        P.PRED_GREATER_THAN: [106, 106],
        P.PRED_GREATER_THAN_EQUALS: [104, 106] * 2,
        P.PRED_IS_NULL: [],
        P.PRED_IS_NOT_NULL: [100, 102, 104, 106] * 2,
    }
    for table in int_int_tables(encoding):
        first = orc.table_scan(table, Predicate(1, P.PRED_LESS_THAN, 108))
        for condition, expected in expectations.items():
            values, _ = scan_values(table, 0, condition, 4, 1, input_filter=first)
            assert values == sorted(expected), (encoding, condition)


@pytest.mark.parametrize("encoding", ENCODINGS)
@pytest.mark.parametrize("value,expectations", [
    (30, {P.PRED_EQUALS: 0, P.PRED_NOT_EQUALS: 14, P.PRED_LESS_THAN: 14, P.PRED_LESS_THAN_EQUALS: 14,
          P.PRED_GREATER_THAN: 0, P.PRED_GREATER_THAN_EQUALS: 0}),      # :486-509
    (-10, {P.PRED_EQUALS: 0, P.PRED_NOT_EQUALS: 14, P.PRED_LESS_THAN: 0, P.PRED_LESS_THAN_EQUALS: 0,
           P.PRED_GREATER_THAN: 14, P.PRED_GREATER_THAN_EQUALS: 14}),   # :511-534
])
def test_scan_value_outside_dictionary(encoding, value, expectations):
    for table in int_int_tables(encoding):
        for condition, count in expectations.items():
            values, _ = scan_values(table, 0, condition, value, 1)
            assert len(values) == count
            if count:
                assert values == sorted([100, 102, 104, 106, 108, 110, 112] * 2)


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_scan_around_bounds(encoding):
    # table_scan_test.cpp:598-621
    rest = [102, 104, 106, 108, 110, 112]
    expectations = {
        P.PRED_EQUALS: [100, 100],
        P.PRED_LESS_THAN: [],
        P.PRED_LESS_THAN_EQUALS: [100, 100],
        P.PRED_GREATER_THAN: rest * 2,
        P.PRED_GREATER_THAN_EQUALS: ([100] + rest) * 2,
        P.PRED_NOT_EQUALS: rest * 2,
        P.PRED_IS_NULL: [],
        P.PRED_IS_NOT_NULL: ([100] + rest) * 2,
    }
    for table in int_int_tables(encoding):
        for condition, expected in expectations.items():
            values, _ = scan_values(table, 0, condition, 0, 1)
            assert values == sorted(expected), (encoding, condition)


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_single_and_double_scan(encoding):
    # table_scan_test.cpp:272-296 -> int_float_filtered2.tbl / int_float_filtered.tbl
    table = tbl("int_float.tbl", 2).encode(encoding)
    first = orc.table_scan(table, Predicate(0, P.PRED_GREATER_THAN_EQUALS, 1234))
    expected = tbl("int_float_filtered2.tbl", 1)
    got = sorted(zip(column_values_at(table, 0, first.row_ids), column_values_at(table, 1, first.row_ids)))
    want = sorted(zip(expected.column_values(0)[0].tolist(), expected.column_values(1)[0].tolist()))
    assert got == want
    # 457.9 is a double literal; lossless_predicate_cast turns `float < 457.9` into `float <= prev(457.9f)`
    bound = np.float32(457.9)
    if float(bound) >= 457.9:
        bound = np.nextafter(bound, np.float32(-np.inf), dtype=np.float32)
    second = orc.table_scan(table, Predicate(1, P.PRED_LESS_THAN_EQUALS, bound), first)
    expected = tbl("int_float_filtered.tbl", 2)
    got = sorted(zip(column_values_at(table, 0, second.row_ids), column_values_at(table, 1, second.row_ids)))
    want = sorted(zip(expected.column_values(0)[0].tolist(), expected.column_values(1)[0].tolist()))
    assert got == want


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_sorted_segment_scans(encoding):
    # table_scan_test.cpp:298-336 (the sorted-search path returns the same rows as the generic path)
    table = tbl("int_sorted.tbl", 4).encode(encoding)
    for condition, expected_file in [(P.PRED_EQUALS, "int_sorted_filtered.tbl"),
                                     (P.PRED_NOT_EQUALS, "int_sorted_filtered2.tbl")]:
        values, _ = scan_values(table, 0, condition, 2, 0)
        assert values == sorted(tbl(expected_file).column_values(0)[0].tolist())
    values, _ = scan_values(table, 0, P.PRED_EQUALS, 6, 0)
    assert values == []


@pytest.mark.parametrize("encoding", ["Unencoded", "Dictionary"])
def test_scan_on_wide_dictionary_segment(encoding):
    # table_scan_test.cpp:636-652: attribute vector width switches u8 -> u16 -> u32; expects 57 and 37 rows
    for entries, threshold, expected in [((1 << 8) + 1, 200, 57), ((1 << 16) + 1, 65500, 37)]:
        table = Table.from_columns([ColumnDefinition("a", capi.TYPE_INT32)],
                                   [np.arange(entries + 1, dtype=np.int32)], chunk_size=100_000).encode(encoding)
        if encoding == "Dictionary":
            width = {257: capi.VEC_FIXED_2B, 65537: capi.VEC_FIXED_4B}[entries]
            assert table.chunks[0].segments[0].vector_type == width
        result = orc.table_scan(table, Predicate(0, P.PRED_GREATER_THAN, threshold))
        assert len(result.row_ids) == expected


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_scan_with_null_values(encoding):
    # table_scan_test.cpp:661-700 (int_float_with_null.tbl): NULLs never match a comparison; IS NULL / IS NOT NULL
    table = tbl("int_float_with_null.tbl", 2).encode(encoding)
    values, nulls = table.column_values(0)
    non_null = sorted(values[~nulls].tolist())
    for condition, reference in [
        (P.PRED_EQUALS, lambda v: v == 1234), (P.PRED_NOT_EQUALS, lambda v: v != 1234),
        (P.PRED_LESS_THAN, lambda v: v < 1234), (P.PRED_LESS_THAN_EQUALS, lambda v: v <= 1234),
        (P.PRED_GREATER_THAN, lambda v: v > 1234), (P.PRED_GREATER_THAN_EQUALS, lambda v: v >= 1234),
    ]:
        got, _ = scan_values(table, 0, condition, 1234, 0)
        assert got == [v for v in non_null if reference(v)]
    result = orc.table_scan(table, Predicate(0, P.PRED_IS_NULL))
    assert len(result.row_ids) == int(nulls.sum())
    result = orc.table_scan(table, Predicate(0, P.PRED_IS_NOT_NULL))
    assert len(result.row_ids) == int((~nulls).sum())


def test_output_is_ascending_per_chunk():
    table = tbl("int_int_shuffled.tbl", 7).encode("Dictionary")
    result = orc.table_scan(table, Predicate(0, P.PRED_GREATER_THAN, 3))
    for chunk_id in range(table.chunk_count):
        rows = result.chunk(chunk_id)
        assert (rows["chunk_id"] == chunk_id).all()
        assert (np.diff(rows["chunk_offset"].astype(np.int64)) > 0).all()


BETWEEN_TYPES = [capi.TYPE_INT32, capi.TYPE_INT64, capi.TYPE_FLOAT32, capi.TYPE_FLOAT64]


@pytest.mark.parametrize("nullable", [False, True])
@pytest.mark.parametrize("sort_mode", [None, "ascending", "descending"])
@pytest.mark.parametrize("encoding", ENCODINGS)
@pytest.mark.parametrize("data_type", BETWEEN_TYPES)
def test_between_reference_expectations(data_type, encoding, sort_mode, nullable):
    # table_scan_between_test.cpp:194-243 on the fixture of :43-96 (string columns excluded: strings reach the scan as
    # value-ID bounds)
    from helpers import BETWEEN_CASES, BETWEEN_CONDITIONS, between_bounds, between_expected, between_fixture

    if encoding == "FrameOfReference" and data_type != capi.TYPE_INT32:
        pytest.skip("FrameOfReference encodes int32 only (the reference's test skips unsupported combinations)")
    table = between_fixture(data_type, encoding, sort_mode, nullable)
    for name, cases in BETWEEN_CASES.items():
        for left, right, expected_with_null in cases:
            lower, upper = between_bounds(data_type, left, right)
            values, _ = scan_values(table, 0, BETWEEN_CONDITIONS[name], lower, 1, upper=upper)
            assert values == between_expected(expected_with_null, sort_mode, nullable), (name, left, right)
