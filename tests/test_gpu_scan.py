"""TableScan parity: the CUDA path (through the C-ABI) against the oracle, bit-exact PosLists."""
import numpy as np
import pytest

import oracle_lib as orc
from helpers import ENCODINGS, assert_pos_lists_equal, random_table, row_ids_equal, tbl
from hyrise_b200 import capi
from hyrise_b200.device import Predicate
from hyrise_b200.storage import ColumnDefinition, Table

pytestmark = pytest.mark.gpu
P = capi


@pytest.fixture(autouse=True, params=["single-pass", "two-pass"])
def scan_variant(device, request):
    """Every scan test runs with both unfiltered scan pipelines: the single-pass ordered compaction with a decoupled look-back
    and the two-pass form (match bits + tile counts, prefix, expansion)."""
    device.set_option("scan_two_pass", "1" if request.param == "two-pass" else "0")
    yield request.param
    device.set_option("scan_two_pass", "0")

BINARY = [P.PRED_EQUALS, P.PRED_NOT_EQUALS, P.PRED_LESS_THAN, P.PRED_LESS_THAN_EQUALS, P.PRED_GREATER_THAN,
          P.PRED_GREATER_THAN_EQUALS]
BETWEEN = [P.PRED_BETWEEN_INCLUSIVE, P.PRED_BETWEEN_LOWER_EXCLUSIVE, P.PRED_BETWEEN_UPPER_EXCLUSIVE,
           P.PRED_BETWEEN_EXCLUSIVE]
NULLS = [P.PRED_IS_NULL, P.PRED_IS_NOT_NULL]


def check_scan(device, table, device_table, predicate, device_filter=None, oracle_filter=None):
    expected = orc.table_scan(table, predicate, oracle_filter)
    result = device.table_scan(device_table, predicate, device_filter)
    try:
        assert_pos_lists_equal(result.to_host(), result.chunk_offsets(), expected)
    except AssertionError as error:
        raise AssertionError(f"{predicate}: {error}") from error
    return result, expected


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_reference_fixtures(device, encoding):
    # the inputs of table_scan_test.cpp:407-621, every comparator, values inside / outside / on the dictionary bounds
    for name, chunk_size in [("int_int_shuffled.tbl", 7), ("int_int_shuffled_2.tbl", 5)]:
        table = tbl(name, chunk_size).encode(encoding, [0, 1])
        device_table = device.upload(table)
        for value in (6, 30, -10, 0, 12, 11):
            for condition in BINARY:
                check_scan(device, table, device_table, Predicate(0, condition, value))
        for condition in NULLS:
            check_scan(device, table, device_table, Predicate(0, condition))
        for condition in BETWEEN:
            for lower, upper in [(2, 8), (0, 12), (-5, 0), (12, 40), (7, 7), (8, 2)]:
                check_scan(device, table, device_table, Predicate(0, condition, lower, upper))
        device_table.drop()


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_null_fixtures(device, encoding):
    for name, chunk_size in [("int_float_with_null.tbl", 2), ("int_int_w_null_8_rows.tbl", 4),
                             ("int_with_nulls_large.tbl", 100)]:
        table = tbl(name, chunk_size).encode(encoding)
        device_table = device.upload(table)
        column_values, _ = table.column_values(0)
        probe_values = sorted(set(column_values.tolist()))[:4] + [1234, 0]
        for value in probe_values:
            for condition in BINARY:
                check_scan(device, table, device_table, Predicate(0, condition, value))
        for condition in NULLS:
            check_scan(device, table, device_table, Predicate(0, condition))
        device_table.drop()


@pytest.mark.parametrize("encoding", ENCODINGS)
@pytest.mark.parametrize("bitpacking", [False, True])
def test_random_all_types(device, encoding, bitpacking):
    rng = np.random.default_rng(1234)
    table = random_table(rng, 25_000, 4_099).encode(encoding, bitpacking=bitpacking)
    device_table = device.upload(table)
    literals = {0: [-1000, -3, 0, 999, 5000], 1: [-(10 ** 12), 17, 10 ** 13], 2: [-1.0, 3.125, 100.0, 1e9],
                3: [-250.0, 0.0, 42.42, float("inf")], 4: [0, 128, 299, 300]}
    for column_id, values in literals.items():
        for value in values:
            for condition in BINARY:
                check_scan(device, table, device_table, Predicate(column_id, condition, value))
        for condition in BETWEEN:
            check_scan(device, table, device_table, Predicate(column_id, condition, values[0], values[-2]))
            check_scan(device, table, device_table, Predicate(column_id, condition, values[1], values[1]))
        for condition in NULLS:
            check_scan(device, table, device_table, Predicate(column_id, condition))
    device_table.drop()


def test_integer_bounds_edge_cases(device):
    int_min, int_max = np.iinfo(np.int32).min, np.iinfo(np.int32).max
    values = np.array([int_min, int_min + 1, -1, 0, 1, int_max - 1, int_max] * 3, dtype=np.int32)
    table = Table.from_columns([ColumnDefinition("a", capi.TYPE_INT32)], [values], chunk_size=8)
    for encoding in ENCODINGS:
        table.encode(encoding)
        device_table = device.upload(table)
        for value in (int_min, int_max, 0):
            for condition in BINARY:
                check_scan(device, table, device_table, Predicate(0, condition, value))
        for condition in BETWEEN:
            check_scan(device, table, device_table, Predicate(0, condition, int_min, int_max))
            check_scan(device, table, device_table, Predicate(0, condition, int_max, int_max))
            check_scan(device, table, device_table, Predicate(0, condition, int_min, int_min))
        device_table.drop()


def test_nan_semantics(device):
    values = np.array([np.nan, 1.0, -np.inf, np.inf, 2.5, np.nan, 0.0, -0.0], dtype=np.float32)
    table = Table.from_columns([ColumnDefinition("f", capi.TYPE_FLOAT32)], [values], chunk_size=5)
    device_table = device.upload(table)
    for value in (1.0, 0.0, float("inf"), float("-inf")):
        for condition in BINARY:
            check_scan(device, table, device_table, Predicate(0, condition, value))
    device_table.drop()


def test_empty_and_ragged_inputs(device):
    # one-row chunks, chunks not divisible by 8, a 4096-row tile boundary, empty result
    for rows, chunk_size in [(1, 1), (9, 1), (4096 * 2 + 3, 4097), (8193, 8193), (30, 7)]:
        values = (np.arange(rows, dtype=np.int32) * 7919) % 101
        table = Table.from_columns([ColumnDefinition("a", capi.TYPE_INT32)], [values], chunk_size=chunk_size)
        for encoding in ENCODINGS:
            table.encode(encoding)
            device_table = device.upload(table)
            for condition in BINARY:
                check_scan(device, table, device_table, Predicate(0, condition, 50))
            check_scan(device, table, device_table, Predicate(0, P.PRED_GREATER_THAN, 1000))  # no match anywhere
            device_table.drop()


def test_string_dictionary_value_id_bounds(device):
    # l_shipdate-style column: 'YYYY-MM-DD' strings, dictionary per chunk, device sees value-IDs only
    rng = np.random.default_rng(7)
    days = rng.integers(0, 2526, 40_000)
    dates = (np.datetime64("1992-01-02") + days).astype("datetime64[D]").astype(str).astype("S10")
    table = Table.from_columns([ColumnDefinition("l_shipdate", capi.TYPE_STRING)], [dates], chunk_size=6_000)
    table.encode("Dictionary")
    assert table.chunks[0].segments[0].vector_type == capi.VEC_FIXED_2B
    device_table = device.upload(table)
    for value in (b"1995-01-01", b"1992-01-01", b"1999-12-31", bytes(dates[17])):
        for condition in BINARY:
            check_scan(device, table, device_table, Predicate(0, condition, value))
    for condition in BETWEEN:
        check_scan(device, table, device_table, Predicate(0, condition, b"1994-01-01", b"1995-01-01"))
    device_table.drop()


@pytest.mark.parametrize("encoding", ENCODINGS)
def test_scan_on_scan_output(device, encoding):
    # reference-table input: TableScan on the output of a TableScan (table_scan_test.cpp:433-463, Q6's scan chain)
    rng = np.random.default_rng(99)
    table = random_table(rng, 30_000, 3_001).encode(encoding)
    device_table = device.upload(table)
    first, first_expected = check_scan(device, table, device_table, Predicate(4, P.PRED_LESS_THAN, 200))
    second, second_expected = check_scan(device, table, device_table, Predicate(0, P.PRED_BETWEEN_INCLUSIVE, -500, 500),
                                         first, first_expected)
    check_scan(device, table, device_table, Predicate(2, P.PRED_GREATER_THAN_EQUALS, 10.0), second, second_expected)
    check_scan(device, table, device_table, Predicate(3, P.PRED_IS_NOT_NULL), second, second_expected)
    device_table.drop()


def test_large_column_properties(device):
    """SF1-sized single column (6 M rows, 65 535-row chunks): size-independent properties + a sampled oracle check."""
    rng = np.random.default_rng(5)
    rows = 6_001_215
    ids = rng.integers(0, 2526, rows).astype(np.int32)
    table = Table.from_columns([ColumnDefinition("d", capi.TYPE_INT32)], [ids]).encode("Dictionary")
    device_table = device.upload(table)
    result = device.table_scan(device_table, Predicate(0, P.PRED_LESS_THAN, 1095))
    rows_out, offsets = result.to_host(), result.chunk_offsets()
    assert len(rows_out) == int((ids < 1095).sum())
    global_index = rows_out["chunk_id"].astype(np.int64) * capi.DEFAULT_CHUNK_SIZE + rows_out["chunk_offset"]
    assert (np.diff(global_index) > 0).all()                      # sorted, no duplicates
    assert np.array_equal(np.flatnonzero(ids < 1095), global_index)  # exactly the matching rows
    per_chunk = np.add.reduceat((ids < 1095).astype(np.int64), np.arange(0, rows, capi.DEFAULT_CHUNK_SIZE))
    assert np.array_equal(np.diff(offsets.astype(np.int64)), per_chunk)
    complement = device.table_scan(device_table, Predicate(0, P.PRED_GREATER_THAN_EQUALS, 1095))
    assert complement.info()[0] + len(rows_out) == rows
    device_table.drop()


@pytest.mark.parametrize("bulk", ["1", "0"])
def test_between_reference_fixture(device, bulk):
    """table_scan_between_test.cpp:194-243 on the fixture of :43-96 (types x encodings x sort modes x nullability), with and
    without the cp.async.bulk staged kernel: the literal expectations of the reference, and the oracle's PosLists."""
    from helpers import BETWEEN_CASES, BETWEEN_CONDITIONS, between_bounds, between_expected, between_fixture, column_values_at

    device.set_option("scan_bulk", bulk)
    try:
        for data_type in (capi.TYPE_INT32, capi.TYPE_INT64, capi.TYPE_FLOAT32, capi.TYPE_FLOAT64):
            for encoding in ENCODINGS:
                if encoding == "FrameOfReference" and data_type != capi.TYPE_INT32:
                    continue
                for sort_mode in (None, "ascending", "descending"):
                    for nullable in (False, True):
                        table = between_fixture(data_type, encoding, sort_mode, nullable)
                        device_table = device.upload(table)
                        for name, cases in BETWEEN_CASES.items():
                            for left, right, expected_with_null in cases:
                                lower, upper = between_bounds(data_type, left, right)
                                predicate = Predicate(0, BETWEEN_CONDITIONS[name], lower, upper)
                                result, _ = check_scan(device, table, device_table, predicate)
                                values = sorted(column_values_at(table, 1, result.to_host()))
                                assert values == between_expected(expected_with_null, sort_mode, nullable), (name, left, right)
                                result.free()
                        device_table.drop()
    finally:
        device.set_option("scan_bulk", "1")
