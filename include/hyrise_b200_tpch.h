/*
 * hyrise_b200_tpch.h — seeded, TPC-H-shaped synthetic workload generator (bench/test tooling, host only).
 *
 * Produces `lineitem` and `orders` with the columns on the hot path, already chunked (65 535 rows, storage/chunk.hpp:52)
 * and encoded the way the reference's "Automatic" encoding leaves them (segment_encoding_utils.cpp:105-115,
 * benchmark_table_encoder.cpp:119-120): int32 -> FrameOfReference, unique int32 (PK) -> Unencoded, everything else ->
 * Dictionary with FixedWidthInteger attribute vectors; dictionaries are per segment (dictionary_encoder.hpp:65-68).
 * Value distributions follow the TPC-H spec constants in third_party/tpch-dbgen/dss.h:331-353.
 * Every value is a pure function of (seed, order index, line number): any chunk can be regenerated independently, and
 * the CPU oracle and the GPU see byte-identical segments.
 *
 * Strings never reach the device. Date columns are 'YYYY-MM-DD' strings in Hyrise; here their per-chunk dictionaries
 * are exposed as sorted day numbers (days since 1992-01-01), whose order equals the ISO string order, so a caller
 * computes DictionarySegment::lower_bound/upper_bound (dictionary_segment.cpp:94-119) with a binary search on them.
 */
#ifndef HYRISE_B200_TPCH_H
#define HYRISE_B200_TPCH_H

#include "hyrise_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hyb_tpch hyb_tpch;

/* lineitem column ids */
enum {
  HYB_L_ORDERKEY = 0,      /* int32   FrameOfReference(u16 offsets) */
  HYB_L_QUANTITY = 1,      /* float   Dictionary u8  (50 values) */
  HYB_L_EXTENDEDPRICE = 2, /* float   Dictionary u16 (~60 K values per chunk) */
  HYB_L_DISCOUNT = 3,      /* float   Dictionary u8  (11 values) */
  HYB_L_TAX = 4,           /* float   Dictionary u8  (9 values) */
  HYB_L_RETURNFLAG = 5,    /* string  Dictionary u8, dictionary_codes = 2 + char (aggregate_hash.cpp:876-878) */
  HYB_L_LINESTATUS = 6,    /* string  Dictionary u8 */
  HYB_L_SHIPDATE = 7,      /* string  Dictionary u16 (<= 2527 dates) */
  HYB_L_COLUMN_COUNT = 8
};
/* orders column ids */
enum {
  HYB_O_ORDERKEY = 0,  /* int32 Unencoded (unique) */
  HYB_O_ORDERDATE = 1, /* string Dictionary u16 */
  HYB_O_COLUMN_COUNT = 2
};

typedef void* (*hyb_tpch_alloc_fn)(size_t bytes);
typedef void (*hyb_tpch_free_fn)(void* ptr);

/*
 * orders = round(1 500 000 * scale_factor) rows; lineitem = 1..7 lines per order (≈ 4x). `alloc`/`free_fn` (may be
 * NULL -> malloc/free) provide the memory of the segment buffers, e.g. pinned memory so uploads run at PCIe speed.
 */
int hyb_tpch_generate(double scale_factor, uint64_t seed, int32_t threads, hyb_tpch_alloc_fn alloc,
                      hyb_tpch_free_fn free_fn, hyb_tpch** out);
/* A shard of a larger data set: the orders with global index first_order + 1 .. first_order + round(1.5 M * scale_factor)
 * and their lineitems (disjoint key ranges across shards; every value depends on the global order index only). */
int hyb_tpch_generate_shard(double scale_factor, uint64_t seed, uint64_t first_order, int32_t threads,
                            hyb_tpch_alloc_fn alloc, hyb_tpch_free_fn free_fn, hyb_tpch** out);
void hyb_tpch_free(hyb_tpch* tables);

int hyb_tpch_lineitem(const hyb_tpch* tables, hyb_table_view* out_view, uint64_t* out_rows);
int hyb_tpch_orders(const hyb_tpch* tables, hyb_table_view* out_view, uint64_t* out_rows);

/* Sorted day numbers of the date dictionary of (table, column, chunk): table 0 = lineitem, 1 = orders. */
int hyb_tpch_date_dictionary(const hyb_tpch* tables, int32_t table, uint32_t column, uint32_t chunk,
                             const int32_t** out_days, uint32_t* out_size);
/* What the operator shim computes per query for a string column: DictionarySegment::lower_bound / upper_bound
 * (dictionary_segment.cpp:94-119) of `value_count` date values on every chunk's dictionary. out_bounds receives
 * chunk_count * 2 * value_count entries laid out [chunk][value][lower, upper] (HYB_INVALID_VALUE_ID past the end) — the
 * layout hyb_scan_predicate::value_id_bounds expects for value_count 1 (binary) and 2 (between). */
int hyb_tpch_value_id_bounds(const hyb_tpch* tables, int32_t table, uint32_t column, const int32_t* day_numbers,
                             uint32_t value_count, uint32_t* out_bounds);
/* Characters of a one-char string dictionary (l_returnflag / l_linestatus). */
int hyb_tpch_char_dictionary(const hyb_tpch* tables, uint32_t column, uint32_t chunk, const char** out_chars,
                             uint32_t* out_size);
/* Total bytes of all segment buffers of a table (what an upload moves over PCIe). */
uint64_t hyb_tpch_table_bytes(const hyb_tpch* tables, int32_t table);

/* The host blocks all segment buffers of both tables live in (for hyb_blocks_upload). out_blocks may be NULL to query
 * the count; bytes = the used part of each block. */
int hyb_tpch_host_blocks(const hyb_tpch* tables, hyb_host_block* out_blocks, uint32_t* out_count);

/* days since 1992-01-01 of 'YYYY-MM-DD' */
int32_t hyb_tpch_day_number(int32_t year, int32_t month, int32_t day);

#ifdef __cplusplus
}
#endif
#endif
