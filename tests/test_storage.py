"""Host-side storage mirror: numpy encoders vs the oracle's C++ restatement vs the reference's known answers
(dictionary_segment_test.cpp, encoded_segment_test.cpp, compressed_vector_test.cpp). CPU only."""
import numpy as np
import pytest

import oracle_lib as orc
from helpers import random_table, tbl
from hyrise_b200 import capi
from hyrise_b200.storage import (ColumnDefinition, Table, compress_bitpacking, compress_fixed_width, decompress_vector,
                                 encode_dictionary, encode_frame_of_reference, short_string_code)


def test_dictionary_is_sorted_unique_and_width_switches():
    # dictionary_segment_test.cpp:47-121 (sorted dictionary, NULL id = size) and :152-182 (u8 -> u16 -> u32 at 2^8 / 2^16)
    values = np.array([6, 4, 3, 4, 5, 6], dtype=np.int32)
    segment = encode_dictionary(values, None, capi.TYPE_INT32)
    assert segment.values.tolist() == [3, 4, 5, 6]
    assert segment.value_ids().tolist() == [3, 1, 0, 1, 2, 3]
    assert segment.vector_type == capi.VEC_FIXED_1B
    nulls = np.array([0, 1, 0, 0, 1, 0], dtype=bool)
    segment = encode_dictionary(values, nulls, capi.TYPE_INT32)
    assert segment.values.tolist() == [3, 4, 6] and segment.dictionary_size == 3
    assert segment.value_ids().tolist() == [2, 3, 0, 1, 3, 2]           # NULL -> value-ID == dictionary size
    for count, expected in [(255, capi.VEC_FIXED_1B), (256, capi.VEC_FIXED_2B), (65535, capi.VEC_FIXED_2B),
                            (65536, capi.VEC_FIXED_4B)]:
        segment = encode_dictionary(np.arange(count, dtype=np.int32), None, capi.TYPE_INT32)
        assert segment.vector_type == expected, count                   # the largest id in use is the NULL id == count


def test_frame_of_reference_known_answers():
    # encoded_segment_test.cpp:676-720: minima per 2048-row block, NULL rows get offset 0, all-NULL block -> INT32_MAX
    rows = 5000
    values = (np.arange(rows, dtype=np.int32) * 3 + 17) % 9973
    nulls = np.zeros(rows, dtype=bool)
    nulls[100:200] = True
    nulls[4096:] = True                                                  # third block entirely NULL
    segment = encode_frame_of_reference(values, nulls)
    assert segment.values.tolist()[:2] == [int(values[:2048][~nulls[:2048]].min()), int(values[2048:4096].min())]
    assert segment.values[2] == np.iinfo(np.int32).max
    offsets = segment.value_ids()
    assert (offsets[100:200] == 0).all()
    decoded = segment.decode()
    assert np.array_equal(decoded[~nulls], values[~nulls])
    minima, oracle_offsets, max_offset, has_nulls = orc.encode_frame_of_reference(values, nulls)
    assert np.array_equal(minima, segment.values) and np.array_equal(oracle_offsets, offsets) and has_nulls
    assert max_offset == int(offsets.max())


@pytest.mark.parametrize("data_type,dtype", [(capi.TYPE_INT32, np.int32), (capi.TYPE_INT64, np.int64),
                                             (capi.TYPE_FLOAT32, np.float32), (capi.TYPE_FLOAT64, np.float64)])
def test_dictionary_encoder_matches_oracle(data_type, dtype):
    rng = np.random.default_rng(1)
    values = rng.integers(-500, 500, 10_000).astype(dtype)
    nulls = rng.random(10_000) < 0.1
    segment = encode_dictionary(values, nulls, data_type)
    dictionary, ids = orc.encode_dictionary(values, nulls, data_type)
    assert np.array_equal(dictionary, segment.values)
    assert np.array_equal(ids, segment.value_ids())
    decoded, decoded_nulls = orc.decode_segment(segment)
    assert np.array_equal(decoded_nulls, nulls) and np.array_equal(decoded[~nulls], values[~nulls])


def test_vector_compression_matches_oracle():
    rng = np.random.default_rng(2)
    for maximum in (1, 2, 5, 200, 255, 256, 4000, 65535, 65536, 10 ** 6, 2 ** 31):
        ids = rng.integers(0, maximum + 1, 3_333).astype(np.uint32)
        ids[7] = maximum
        ours, vector_type = compress_fixed_width(ids, maximum)
        theirs, oracle_type = orc.compress_fixed_width(ids, maximum)
        assert vector_type == oracle_type and np.array_equal(ours, theirs)
        words, bits = compress_bitpacking(ids)
        oracle_words, oracle_bits = orc.compress_bitpacking(ids)
        assert bits == oracle_bits == max(1, int(np.ceil(np.log2(maximum + 1))))  # bitpacking_compressor.cpp:24-32
        assert np.array_equal(words, oracle_words)
        assert np.array_equal(decompress_vector(words, capi.VEC_BITPACKED, bits, len(ids)), ids)
    words, bits = compress_bitpacking(np.zeros(10, dtype=np.uint32))
    assert bits == 1                                                     # all zeros still need one bit


def test_oracle_decodes_every_encoding():
    rng = np.random.default_rng(3)
    table = random_table(rng, 9_000, 2_049)
    expected = [table.column_values(c) for c in range(table.column_count)]
    for encoding in ("Dictionary", "FrameOfReference", "Unencoded"):
        for bitpacking in (False, True):
            table.encode(encoding, bitpacking=bitpacking)
            for column in range(table.column_count):
                values, nulls = table.column_values(column)
                assert np.array_equal(nulls, expected[column][1])
                assert np.array_equal(values[~nulls], expected[column][0][~nulls])
                chunk_values = np.concatenate([orc.decode_segment(chunk.segments[column])[0] for chunk in table.chunks])
                assert np.array_equal(chunk_values[~nulls], expected[column][0][~nulls])


def test_short_string_keys():
    # aggregate_hash.cpp:863-899
    assert short_string_code(b"") == 1
    assert short_string_code(b"A") == 2 + 65
    assert short_string_code(b"\xff") == 257
    assert short_string_code(b"ab") == 258 + 97 + (98 << 8)
    assert short_string_code(b"abcd") == 16_843_010 + 97 + (98 << 8) + (99 << 16) + (100 << 24)
    assert short_string_code(b"abcde") is None


def test_tbl_loader():
    table = tbl("int_float_with_null.tbl", 2)
    assert [d.name for d in table.column_definitions] == ["a", "b"]
    assert table.column_definitions[0].nullable and table.column_definitions[1].data_type == capi.TYPE_FLOAT32
    values, nulls = table.column_values(0)
    assert values[0] == 12345 and nulls.tolist()[:3] == [False, False, True]
    assert tbl("join_test_runner/input_table_left_0.tbl").row_count == 0


def test_generator_matches_reference_encoders():
    """The C++ generator's segments decode to values whose re-encoding with the numpy encoders gives the same bytes."""
    from hyrise_b200.tpch import TpchTables, L_EXTENDEDPRICE, L_ORDERKEY, L_QUANTITY

    tables = TpchTables(0.02, seed=7)
    lineitem = tables.lineitem
    assert lineitem.chunk_count == 2
    import ctypes as C

    for column, numpy_type in ((L_QUANTITY, np.float32), (L_EXTENDEDPRICE, np.float32)):
        desc = lineitem.segment_desc(0, column)
        dictionary = np.ctypeslib.as_array(C.cast(desc.values, C.POINTER(C.c_float)), shape=(desc.dictionary_size,)).copy()
        assert (np.diff(dictionary) > 0).all()                           # sorted, unique
        width = {capi.VEC_FIXED_1B: C.c_uint8, capi.VEC_FIXED_2B: C.c_uint16, capi.VEC_FIXED_4B: C.c_uint32}[desc.vector_type]
        ids = np.ctypeslib.as_array(C.cast(desc.attribute_vector, C.POINTER(width)), shape=(desc.row_count,)).astype(np.uint32)
        again = encode_dictionary(dictionary[ids], None, capi.TYPE_FLOAT32)
        assert np.array_equal(again.values, dictionary) and np.array_equal(again.value_ids(), ids)
        assert again.vector_type == desc.vector_type
    desc = lineitem.segment_desc(0, L_ORDERKEY)
    assert desc.encoding == capi.ENC_FRAME_OF_REFERENCE and desc.vector_type == capi.VEC_FIXED_2B
    tables.close()
