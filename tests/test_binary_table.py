"""The loading path: Hyrise's binary table format -> host segments in the device pool's layout.
`hyrise_b200/binary_table.py` against the reference's own fixtures (resources/test_data/bin, copied by
tests/golden/make_golden.py) and the tables `src/test/lib/import_export/binary/binary_parser_test.cpp` expects for them."""
import os

import numpy as np
import pytest

from hyrise_b200 import capi
from hyrise_b200.binary_table import BinaryFormatError, read_binary_table, write_binary_table
from hyrise_b200.storage import ColumnDefinition, Table

BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bin")
ENCODINGS = ["Unencoded", "Dictionary", "RunLength"]   # BinaryParserMultiEncodingTest minus LZ4 (binary_parser_test.cpp:24-27)
NULL = None


def rows_of(table: Table):
    """Row tuples in table order, None for NULL, bytes for strings."""
    out = []
    for chunk in table.chunks:
        columns = []
        for segment in chunk.segments:
            values, nulls = segment.decode(), segment.null_mask()
            columns.append([None if null else (bytes(value) if isinstance(value, bytes) else value.item())
                            for value, null in zip(values, nulls)])
        out.extend(zip(*columns))
    return out


def assert_rows(got, want):
    assert len(got) == len(want)
    for got_row, want_row in zip(got, want):
        for a, b in zip(got_row, want_row):
            if isinstance(b, float):
                assert a == pytest.approx(b, rel=1e-6)
            else:
                assert a == b


ALL_TYPES_SORTED = [(b"AAAAA", 1, 100, 1.1, 11.1), (b"BBBBBBBBBB", 2, 200, 2.2, 22.2), (b"CCCCCCCCCCCCCCC", 3, 300, 3.3, 33.3),
                    (b"DDDDDDDDDDDDDDDDDDDD", 4, 400, 4.4, 44.4)]
# binary_parser_test.cpp:28-283: fixture directory -> (expected chunk count or None, expected rows)
MULTI_ENCODING_CASES = {
    "SingleChunkSingleFloatColumn": (1, [(5.5,), (13.0,), (16.2,)]),
    "MultipleChunkSingleFloatColumn": (2, [(5.5,), (13.0,), (16.2,)]),
    "StringSegment": (2, [(b"This",), (b"is",), (b"a",), (b"test",)]),
    "AllTypesSegmentSorted": (2, ALL_TYPES_SORTED),
    "AllTypesSegmentUnsorted": (2, [ALL_TYPES_SORTED[3], ALL_TYPES_SORTED[0], ALL_TYPES_SORTED[2], ALL_TYPES_SORTED[1]]),
    "AllTypesMixColumn": (2, ALL_TYPES_SORTED),
    "EmptyStringsSegment": (1, [(b"",)] * 5),
    "AllTypesNullValues": (1, [(NULL, 1.1, 100, b"one", 1.11), (2, NULL, 200, b"two", 2.22), (3, 3.3, NULL, b"three", 3.33),
                               (4, 4.4, 400, NULL, 4.44), (5, 5.5, 500, b"five", NULL)]),
    "AllTypesAllNullValues": (1, [(NULL,) * 5] * 5),
    "RepeatedInt": (2, [(1,), (2,), (2,), (2,), (2,), (1,)]),
    "RunNullValues": (1, [(NULL,), (1,), (NULL,), (NULL,), (NULL,), (2,), (2,), (NULL,)]),
}


@pytest.mark.parametrize("encoding", ENCODINGS)
@pytest.mark.parametrize("name", sorted(MULTI_ENCODING_CASES))
def test_multi_encoding_fixtures(name, encoding):
    chunk_count, want = MULTI_ENCODING_CASES[name]
    parsed = read_binary_table(os.path.join(BIN, name, encoding + ".bin"))
    assert parsed.table.chunk_count == chunk_count
    assert_rows(rows_of(parsed.table), want)
    for chunk in parsed.table.chunks:   # everything is in an encoding the device pool uploads, in aligned arena slots
        for segment in chunk.segments:
            assert segment.encoding in (capi.ENC_UNENCODED, capi.ENC_DICTIONARY, capi.ENC_FRAME_OF_REFERENCE)
            for buffer in (segment.values, segment.nulls, segment.attribute_vector, segment.dictionary_codes):
                if buffer is not None and buffer.nbytes:
                    assert buffer.ctypes.data % 256 == 0


def test_lz4_is_reported_as_unsupported():
    with pytest.raises(capi.UnsupportedOnDevice):
        read_binary_table(os.path.join(BIN, "AllTypesSegmentSorted", "LZ4.bin"))


def test_fixed_string_dictionary_fixtures():   # binary_parser_test.cpp:305-351
    words = [(b"This",), (b"is",), (b"a",), (b"test",)]
    assert_rows(rows_of(read_binary_table(os.path.join(BIN, "FixedStringDictionarySingleChunk.bin")).table), words)
    parsed = read_binary_table(os.path.join(BIN, "FixedStringDictionaryMultipleChunks.bin"))
    assert parsed.table.chunk_count == 2
    assert_rows(rows_of(parsed.table), words)
    assert_rows(rows_of(read_binary_table(os.path.join(BIN, "FixedStringDictionaryNullValue.bin")).table),
                [(b"This",), (b"is",), (b"a",), (NULL,), (b"test",), (NULL,)])


def test_frame_of_reference_fixtures():   # binary_parser_test.cpp:353-385 (+ the two fixtures of the writer's tests)
    parsed = read_binary_table(os.path.join(BIN, "NullValuesFrameOfReferenceSegment.bin"))
    assert all(chunk.segments[0].encoding == capi.ENC_FRAME_OF_REFERENCE for chunk in parsed.table.chunks)
    assert_rows(rows_of(parsed.table), [(1,), (NULL,), (2,), (NULL,), (5,)])
    assert_rows(rows_of(read_binary_table(os.path.join(BIN, "AllNullFrameOfReferenceSegment.bin")).table), [(NULL,)] * 5)
    for name in ("SingleChunkFrameOfReferenceSegment.bin", "MultipleChunksFrameOfReferenceSegment.bin"):
        table = read_binary_table(os.path.join(BIN, name)).table
        assert all(chunk.segments[0].encoding == capi.ENC_FRAME_OF_REFERENCE for chunk in table.chunks)
        assert table.row_count > 0


def test_invalid_files_raise():   # binary_parser_test.cpp:387-399
    with pytest.raises(BinaryFormatError):
        read_binary_table(os.path.join(BIN, "InvalidEncodingType.bin"))
    with pytest.raises(BinaryFormatError):
        read_binary_table(os.path.join(BIN, "InvalidAttributeVectorWidth.bin"))
    with pytest.raises(OSError):
        read_binary_table("not_existing_file")


def test_empty_table_and_sort_definitions():   # binary_parser_test.cpp:401-446
    parsed = read_binary_table(os.path.join(BIN, "TwoColumnsNoValues.bin"))
    assert [d.name for d in parsed.table.column_definitions] == ["FirstColumn", "SecondColumn"]
    assert [d.data_type for d in parsed.table.column_definitions] == [capi.TYPE_INT32, capi.TYPE_STRING]
    assert parsed.table.row_count == 0 and parsed.table.target_chunk_size == 30_000
    parsed = read_binary_table(os.path.join(BIN, "SortColumnDefinitions.bin"))
    assert_rows(rows_of(parsed.table), [(1, 3), (2, 2), (3, 1), (1, 3), (2, 2), (1, 1), (1, 1), (2, 2), (1, 1)])
    descending_nulls_first = 1   # SortMode {AscendingNullsFirst, DescendingNullsFirst, AscendingNullsLast, DescendingNullsLast} (types.hpp:219)
    assert parsed.sorted_columns[0][0][0] == 0 and parsed.sorted_columns[0][1] == (1, descending_nulls_first)
    assert parsed.sorted_columns[1] == [(1, descending_nulls_first)] and parsed.sorted_columns[2] == []


def all_fixture_files():
    out = []
    for root, _, files in os.walk(BIN):
        out.extend(os.path.join(root, name) for name in sorted(files) if name.endswith(".bin"))
    return sorted(out)


def assert_same_tables(native, python):
    assert [(d.name, d.data_type, d.nullable) for d in native.column_definitions] == \
           [(d.name, d.data_type, d.nullable) for d in python.column_definitions]
    assert native.target_chunk_size == python.target_chunk_size and native.chunk_count == python.chunk_count
    for a_chunk, b_chunk in zip(native.chunks, python.chunks):
        for a, b in zip(a_chunk.segments, b_chunk.segments):
            assert (a.encoding, a.data_type, a.row_count, a.vector_type, a.bit_width, a.dictionary_size) == \
                   (b.encoding, b.data_type, b.row_count, b.vector_type, b.bit_width, b.dictionary_size)
            assert np.array_equal(a.null_mask(), b.null_mask())
            mask = ~a.null_mask()
            assert np.array_equal(a.decode()[mask], b.decode()[mask])
            for x, y in ((a.attribute_vector, b.attribute_vector), (a.dictionary_codes, b.dictionary_codes)):
                assert (x is None) == (y is None)
                if x is not None:
                    assert np.array_equal(np.asarray(x), np.asarray(y))
            for buffer in (a.values, a.nulls, a.attribute_vector, a.dictionary_codes):
                if buffer is not None and buffer.nbytes:
                    assert buffer.ctypes.data % 256 == 0


@pytest.mark.parametrize("path", all_fixture_files(), ids=lambda p: os.path.relpath(p, BIN))
def test_native_loader_matches_the_python_loader(path):
    """hyb_binary_table_open (csrc/binary_loader.cu) against binary_table.read_binary_table on every reference fixture: same
    definitions, encodings, vectors, dictionaries, group-by codes and sort information — or the same kind of failure."""
    from hyrise_b200.binary_table import NativeBinaryTable

    try:
        python = read_binary_table(path)
    except capi.UnsupportedOnDevice:
        with pytest.raises(capi.UnsupportedOnDevice):
            NativeBinaryTable(path)
        return
    except BinaryFormatError:
        with pytest.raises(capi.HyriseB200Error) as error:
            NativeBinaryTable(path)
        assert error.value.status == capi.HYB_ERR_INVALID
        return
    native = NativeBinaryTable(path)
    assert_same_tables(native.table, python.table)
    assert native.sorted_columns == python.sorted_columns
    native.close()


def test_native_loader_errors_and_string_bounds(tmp_path):
    from hyrise_b200.binary_table import NativeBinaryTable
    from hyrise_b200.device import Predicate
    from test_oracle_aggregate import lineitem_table

    with pytest.raises(capi.HyriseB200Error) as error:
        NativeBinaryTable("not_existing_file")
    assert error.value.status == capi.HYB_ERR_NOT_FOUND
    source, _ = lineitem_table()   # dbgen sf-0.01 lineitem, reference default encodings, strings + FoR + dictionaries
    path = str(tmp_path / "lineitem.bin")
    write_binary_table(source, path)
    native = NativeBinaryTable(path)
    assert_same_tables(native.table, read_binary_table(path).table)
    assert len(native.host_blocks()) >= 1
    for predicate in (Predicate(7, capi.PRED_LESS_THAN, b"1995-01-01"),
                      Predicate(7, capi.PRED_BETWEEN_INCLUSIVE, b"1994-01-01", b"1994-12-31"), Predicate(5, capi.PRED_EQUALS, b"R"),
                      Predicate(7, capi.PRED_GREATER_THAN, b"2999-01-01"), Predicate(7, capi.PRED_GREATER_THAN, b"")):
        between = predicate.upper is not None
        got = native.value_id_bounds(predicate.column_id, predicate.lower, predicate.upper if between else None)
        assert np.array_equal(got, source.string_value_id_bounds(predicate))
    native.close()


@pytest.mark.parametrize("bitpacking", [False, True])
def test_write_then_read_keeps_segments_bit_identical(tmp_path, bitpacking):
    rng = np.random.default_rng(5)
    rows = 10_000
    definitions = [ColumnDefinition("k", capi.TYPE_INT32, True), ColumnDefinition("l", capi.TYPE_INT64), ColumnDefinition("f", capi.TYPE_FLOAT32, True),
                   ColumnDefinition("d", capi.TYPE_FLOAT64), ColumnDefinition("s", capi.TYPE_STRING, True)]
    columns = [rng.integers(-5000, 5000, rows, dtype=np.int32), rng.integers(-10 ** 12, 10 ** 12, rows, dtype=np.int64),
               (rng.integers(0, 500, rows) / 4).astype(np.float32), rng.normal(0, 10, rows),
               np.array([b"v%03d" % v for v in rng.integers(0, 300, rows)], dtype="S8")]
    nulls = [rng.random(rows) < 0.1, None, rng.random(rows) < 0.05, None, rng.random(rows) < 0.2]
    for encoding in ("Unencoded", "Dictionary", "Automatic"):
        table = Table.from_columns(definitions, columns, nulls, chunk_size=3_000).encode(encoding, bitpacking=bitpacking)
        path = str(tmp_path / f"{encoding}.bin")
        write_binary_table(table, path)
        loaded = read_binary_table(path).table
        assert loaded.chunk_count == table.chunk_count and loaded.target_chunk_size == 3_000
        assert [(d.name, d.data_type, d.nullable) for d in loaded.column_definitions] == [(d.name, d.data_type, d.nullable) for d in definitions]
        for before, after in zip(table.chunks, loaded.chunks):
            for a, b in zip(before.segments, after.segments):
                assert (a.encoding, a.vector_type, a.bit_width, a.dictionary_size, a.row_count) == \
                       (b.encoding, b.vector_type, b.bit_width, b.dictionary_size, b.row_count)
                assert np.array_equal(a.null_mask(), b.null_mask())
                mask = ~a.null_mask()
                assert np.array_equal(a.decode()[mask], b.decode()[mask])
                if a.attribute_vector is not None:
                    assert np.array_equal(np.asarray(a.attribute_vector).view(np.uint8)[: b.attribute_vector.nbytes],
                                          np.asarray(b.attribute_vector).view(np.uint8))


@pytest.mark.gpu
def test_binary_table_to_device_through_pinned_blocks(device, tmp_path):
    """binary file -> pinned host blocks -> hyb_blocks_upload (one DMA per block) -> operators; checked against the oracle on
    the table the file was written from (dbgen sf-0.01 lineitem in the reference's default encodings)."""
    import oracle_lib as orc
    from helpers import assert_aggregate_outputs_equal, assert_pos_lists_equal
    from hyrise_b200.device import Predicate
    from test_oracle_aggregate import Q1_AGGREGATES, Q1_PREDICATES, lineitem_table

    source, _ = lineitem_table()
    path = str(tmp_path / "lineitem.bin")
    write_binary_table(source, path)
    parsed = read_binary_table(path, context=device)
    block_set = device.upload_blocks(parsed.host_blocks())
    table = device.upload_from_blocks(parsed.table, block_set)
    predicate = Predicate(1, capi.PRED_LESS_THAN, 24.0)
    result = device.table_scan(table, predicate)
    assert_pos_lists_equal(result.to_host(), result.chunk_offsets(), orc.table_scan(source, predicate))
    result.free()
    got = device.aggregate_hash(table, [5, 6], Q1_AGGREGATES, predicates=Q1_PREDICATES)
    assert_aggregate_outputs_equal(got, orc.aggregate_hash(source, [5, 6], Q1_AGGREGATES, predicates=Q1_PREDICATES))
    table.drop()
    device.free_blocks(block_set)
    # the same file through the native loader: parse into pinned blocks and upload in one call (hyb_table_upload_binary)
    from hyrise_b200.binary_table import NativeBinaryTable
    native = NativeBinaryTable(path, pinned=True)
    table = native.upload(device)
    result = device.table_scan(table, predicate)
    assert_pos_lists_equal(result.to_host(), result.chunk_offsets(), orc.table_scan(source, predicate))
    result.free()
    got = device.aggregate_hash(table, [5, 6], Q1_AGGREGATES, predicates=Q1_PREDICATES)
    assert_aggregate_outputs_equal(got, orc.aggregate_hash(source, [5, 6], Q1_AGGREGATES, predicates=Q1_PREDICATES))
    table.drop()
    native.close()
