/*
 * oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. This is a CPU restatement of the reference algorithms (hyrise/hyrise @ 2f7bedf3) for the
 * hot path: segment encoders, TableScan, JoinHash, AggregateHash. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it — and there only as the checker or as the timed CPU baseline. The
 * product (hyrise_b200/, libhyrise_b200.so) never links, imports or calls anything in this directory.
 *
 * Parity pinning: the restatement is checked (tests/test_oracle_*.py) against the reference's own golden vectors —
 * resources/test_data/tbl fixtures and the literal expectations of src/test/lib/operators/{table_scan,join_hash/..,
 * aggregate}_test.cpp and src/test/lib/storage/{dictionary_segment,encoded_segment}_test.cpp — committed under
 * tests/golden/ by tests/golden/make_golden.py. The reference itself cannot be compiled here (needs Boost >= 1.81,
 * oneTBB, sqlite3 headers), so there is no oracle/_ref for the operators.
 */
#ifndef HYRISE_ORACLE_H
#define HYRISE_ORACLE_H

#include "../include/hyrise_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- segment encoders ------------------------------------------------------------------------------------------- */

/* dictionary_encoder.hpp:33-103. out_dictionary: capacity n elements of the type; out_value_ids: n. */
int orc_encode_dictionary(int32_t data_type, const void* values, const uint8_t* nulls, uint32_t n, void* out_dictionary,
                          uint32_t* out_dictionary_size, uint32_t* out_value_ids);
/* fixed_width_integer_compressor.cpp:18-57. out: capacity 4*n bytes. */
int orc_compress_fixed_width(const uint32_t* in, uint32_t n, uint32_t max_value, void* out, int32_t* out_vector_type);
/* bitpacking_compressor.cpp:21-53 + compact_iterator.hpp:218-252. out_words: capacity ceil(n*32/64) words. */
int orc_compress_bitpacking(const uint32_t* in, uint32_t n, uint64_t* out_words, int32_t* out_bit_width);
/* frame_of_reference_encoder.hpp:25-122. out_minima: ceil(n/2048); out_offsets: n. */
int orc_encode_frame_of_reference(const int32_t* values, const uint8_t* nulls, uint32_t n, int32_t* out_minima,
                                  uint32_t* out_offsets, uint32_t* out_max_offset, int32_t* out_contains_nulls);
/* Decode any supported segment back to values (8 bytes stride for 8-byte types, 4 otherwise) + null bytes. For STRING
 * dictionary segments out_values receives the value-IDs as uint32. */
int orc_decode_segment(const hyb_segment_desc* segment, void* out_values, uint8_t* out_nulls);

/* ---- operators -------------------------------------------------------------------------------------------------- */

typedef struct orc_pos_list {
  uint32_t chunk_count;
  uint64_t total;
  uint64_t* chunk_offsets; /* chunk_count + 1 */
  hyb_row_id* row_ids;     /* total */
} orc_pos_list;

/* TableScan over all chunks (table_scan.cpp:97-240). threads <= 1: sequential; else one job per chunk (>= 500 rows,
 * table_scan.cpp:223-229) on a pool of `threads` std::threads. */
int orc_table_scan(const hyb_table_view* table, const hyb_scan_predicate* predicate, const orc_pos_list* input_filter,
                   int32_t threads, orc_pos_list* out);
void orc_pos_list_free(orc_pos_list* list);

typedef struct orc_join_result {
  uint64_t pair_count;
  hyb_row_id* build_row_ids; /* NULL for Semi/Anti */
  hyb_row_id* probe_row_ids;
  int32_t radix_bits;
  uint32_t partition_count;
  uint64_t* partition_offsets; /* partition_count + 1 */
  uint32_t slice_count;        /* pos lists produced by probe(): one per (partition, 131070-row slice) */
  uint64_t* slice_offsets;     /* slice_count + 1 */
  uint32_t output_chunk_count; /* after write_output_chunks merging (join_output_writing.cpp:255-296) */
  uint64_t* output_chunk_offsets; /* output_chunk_count + 1 */
  uint64_t build_materialized;  /* performance_data.build_side_materialized_value_count */
  uint64_t probe_materialized;
} orc_join_result;

/* JoinHash (join_hash.cpp:270-572, join_hash_steps.hpp). radix_bits < 0: calculate_radix_bits (join_hash.cpp:70-114). */
int orc_join_hash(const hyb_table_view* build_table, uint32_t build_column, const orc_pos_list* build_filter,
                  const hyb_table_view* probe_table, uint32_t probe_column, const orc_pos_list* probe_filter,
                  int32_t mode, int32_t radix_bits, int32_t threads, orc_join_result* out);
void orc_join_result_free(orc_join_result* result);
/* Test hook: materialize_input<int32,int32> internals (elements, per-chunk radix histograms, Bloom filter slots). */
int orc_debug_materialize(const hyb_table_view* table, uint32_t column, int32_t keep_nulls, int32_t radix_bits,
                          const uint32_t* input_bloom_slots, uint32_t input_bloom_slot_count, int32_t* out_values,
                          hyb_row_id* out_row_ids, uint8_t* out_nulls, uint64_t* out_count, uint64_t* out_histograms,
                          uint32_t* out_bloom_slots, uint32_t* out_bloom_slot_count);
int32_t orc_calculate_radix_bits(uint64_t build_side_size, uint64_t probe_side_size);

typedef struct orc_aggregate_column {
  int32_t value_type; /* hyb_data_type */
  void* values;       /* group_count elements */
  uint8_t* nulls;     /* group_count */
} orc_aggregate_column;

typedef struct orc_aggregate_result {
  uint64_t group_count;
  int32_t used_immediate_keys;
  hyb_row_id* row_ids;
  uint32_t aggregate_count;
  orc_aggregate_column* columns;
} orc_aggregate_result;

/* [TableScan ...] -> [Projection] -> AggregateHash (aggregate_hash.cpp). query->table / query->filter handles are
 * ignored; the inputs are `table` and `filter`. `parallel` != 0 selects a chunk-parallel variant (the reference's
 * aggregation phase is sequential; the parallel variant is reported separately by bench.py). */
int orc_aggregate_hash(const hyb_table_view* table, const hyb_aggregate_query* query, const orc_pos_list* filter,
                       int32_t threads, int32_t parallel, orc_aggregate_result* out);
void orc_aggregate_result_free(orc_aggregate_result* result);

/* ---- predicate normalisation (lossless_predicate_cast.hpp/.cpp, lossless_cast.hpp, types.cpp, table_scan.cpp:340-441) -- */
int orc_next_float_towards(double value, double towards, float* out_value); /* 1 = has a value */
int orc_normalize_predicate(int32_t condition, int32_t literal_type, hyb_value literal, int32_t column_type,
                            int32_t value_on_left, int32_t* out_condition, hyb_value* out_value);
int orc_normalize_between(int32_t condition, int32_t lower_type, hyb_value lower, int32_t upper_type, hyb_value upper,
                          int32_t column_type, int32_t* out_condition, hyb_value* out_lower, hyb_value* out_upper);

const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
