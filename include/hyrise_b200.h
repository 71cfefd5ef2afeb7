/*
 * hyrise_b200.h — C-ABI of the Blackwell (sm_100a) execution path for Hyrise's three bandwidth-bound operators.
 *
 * This header is the drop-in boundary. The reference (hyrise/hyrise @ 2f7bedf3) has no FFI for operators; the seam is
 * the C++ virtual `AbstractOperator::_on_execute()` (src/lib/operators/abstract_operator.hpp:231). Each entry point
 * below replaces the body of one reference function and is what a maintainer's shim inside that function would call
 * (see INTEGRATION.md for the shim code):
 *
 *   hyb_table_scan        <-  TableScan::_on_execute                 src/lib/operators/table_scan.cpp:97-240
 *                             ColumnVsValueTableScanImpl             .../table_scan/column_vs_value_table_scan_impl.cpp:43-272
 *                             ColumnBetweenTableScanImpl             .../table_scan/column_between_table_scan_impl.cpp:42-226
 *                             ColumnIsNullTableScanImpl              .../table_scan/column_is_null_table_scan_impl.cpp
 *                             AbstractTableScanImpl::_scan_with_iterators  .../table_scan/abstract_table_scan_impl.hpp:56-242
 *   hyb_join_hash         <-  JoinHash::_on_execute / JoinHashImpl   src/lib/operators/join_hash.cpp:116-572
 *                             materialize_input/partition_by_radix/build/probe/probe_semi_anti
 *                                                                    src/lib/operators/join_hash/join_hash_steps.hpp:274-922
 *   hyb_aggregate_hash    <-  AggregateHash::_on_execute/_aggregate  src/lib/operators/aggregate_hash.cpp:950-1372
 *   hyb_table_* (pool)    <-  new "device column pool": device copies of ValueSegment / DictionarySegment /
 *                             FrameOfReferenceSegment               src/lib/storage/{value_segment.hpp:84-85,
 *                             dictionary_segment.hpp:88-90, frame_of_reference_segment.hpp:49,94-97}
 *
 * Conventions:
 *   - plain C: pointers + sizes, no C++/torch types; every function returns hyb_status (0 = OK);
 *     hyb_last_error() returns a thread-local message. No exception crosses this boundary.
 *   - every entry point is thread-safe and re-entrant; inputs are borrowed and immutable for the duration of the call.
 *   - HYB_ERR_UNSUPPORTED means "run the CPU operator" (e.g. LIKE scans, string join keys); anything else is a
 *     hard failure the shim turns into Fail(hyb_last_error()).
 *   - results stay device-resident behind handles (so the next operator can consume them without a PCIe round trip)
 *     and are copied out on request into caller-owned host memory (pinned via hyb_host_alloc for full PCIe speed).
 */
#ifndef HYRISE_B200_H
#define HYRISE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYB_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------------------------------
 * Status codes
 * ---------------------------------------------------------------------------------------------------------------- */
typedef enum hyb_status {
  HYB_OK = 0,
  HYB_ERR_INVALID = 1,     /* bad argument (maps to reference Assert/Fail -> std::logic_error, utils/assert.hpp:48-82) */
  HYB_ERR_UNSUPPORTED = 2, /* not on the GPU path: caller runs the CPU operator */
  HYB_ERR_CUDA = 3,        /* CUDA runtime failure; message holds cudaGetErrorString */
  HYB_ERR_OOM = 4,
  HYB_ERR_NOT_FOUND = 5    /* unknown handle */
} hyb_status;

/* ------------------------------------------------------------------------------------------------------------------
 * Core types — restated from src/lib/types.hpp and src/lib/all_type_variant.hpp:34-39
 * ---------------------------------------------------------------------------------------------------------------- */

/* RowID {ChunkID, ChunkOffset}, 8 bytes — types.hpp:97-117. NULL_ROW_ID = {0xFFFFFFFF, 0xFFFFFFFF}. */
typedef struct hyb_row_id {
  uint32_t chunk_id;
  uint32_t chunk_offset;
} hyb_row_id;

#define HYB_INVALID_CHUNK_ID 0xFFFFFFFFu
#define HYB_INVALID_CHUNK_OFFSET 0xFFFFFFFFu
#define HYB_INVALID_VALUE_ID 0xFFFFFFFFu /* types.hpp INVALID_VALUE_ID */
#define HYB_DEFAULT_CHUNK_SIZE 65535u    /* storage/chunk.hpp:52 */
#define HYB_FOR_BLOCK_SIZE 2048u         /* storage/frame_of_reference_segment.hpp:49 */

typedef enum hyb_data_type { /* all_type_variant.hpp:34-39 (Null excluded) */
  HYB_TYPE_INT32 = 0,
  HYB_TYPE_INT64 = 1,
  HYB_TYPE_FLOAT32 = 2,
  HYB_TYPE_FLOAT64 = 3,
  HYB_TYPE_STRING = 4 /* only value-IDs of string dictionaries reach the device */
} hyb_data_type;

typedef enum hyb_encoding { /* storage/encoding_type.hpp; the three the default TPC-H encoding produces */
  HYB_ENC_UNENCODED = 0,    /* ValueSegment<T> */
  HYB_ENC_DICTIONARY = 1,   /* DictionarySegment<T> */
  HYB_ENC_FRAME_OF_REFERENCE = 2 /* FrameOfReferenceSegment<int32_t> */
} hyb_encoding;

typedef enum hyb_vector_type { /* storage/vector_compression/compressed_vector_type.hpp */
  HYB_VEC_NONE = 0,
  HYB_VEC_FIXED_1B = 1, /* FixedWidthIntegerVector<uint8_t>  */
  HYB_VEC_FIXED_2B = 2, /* FixedWidthIntegerVector<uint16_t> */
  HYB_VEC_FIXED_4B = 3, /* FixedWidthIntegerVector<uint32_t> */
  HYB_VEC_BITPACKED = 4 /* BitPackingVector: LSB-first b-bit fields in uint64 words (compact_iterator.hpp:218-252) */
} hyb_vector_type;

/* PredicateCondition — same order/values as types.hpp:160-179 so a shim can static_cast. */
typedef enum hyb_predicate_condition {
  HYB_PRED_EQUALS = 0,
  HYB_PRED_NOT_EQUALS = 1,
  HYB_PRED_LESS_THAN = 2,
  HYB_PRED_LESS_THAN_EQUALS = 3,
  HYB_PRED_GREATER_THAN = 4,
  HYB_PRED_GREATER_THAN_EQUALS = 5,
  HYB_PRED_BETWEEN_INCLUSIVE = 6,
  HYB_PRED_BETWEEN_LOWER_EXCLUSIVE = 7,
  HYB_PRED_BETWEEN_UPPER_EXCLUSIVE = 8,
  HYB_PRED_BETWEEN_EXCLUSIVE = 9,
  HYB_PRED_IN = 10,                   /* unsupported */
  HYB_PRED_NOT_IN = 11,               /* unsupported */
  HYB_PRED_LIKE = 12,                 /* unsupported */
  HYB_PRED_NOT_LIKE = 13,             /* unsupported */
  HYB_PRED_LIKE_INSENSITIVE = 14,     /* unsupported */
  HYB_PRED_NOT_LIKE_INSENSITIVE = 15, /* unsupported */
  HYB_PRED_IS_NULL = 16,
  HYB_PRED_IS_NOT_NULL = 17
} hyb_predicate_condition;

/* JoinMode — same values as types.hpp:210. */
typedef enum hyb_join_mode {
  HYB_JOIN_INNER = 0,
  HYB_JOIN_LEFT = 1,
  HYB_JOIN_RIGHT = 2,
  HYB_JOIN_FULL_OUTER = 3, /* unsupported (as in JoinHash) */
  HYB_JOIN_CROSS = 4,      /* unsupported */
  HYB_JOIN_SEMI = 5,
  HYB_JOIN_ANTI_NULL_AS_TRUE = 6,
  HYB_JOIN_ANTI_NULL_AS_FALSE = 7
} hyb_join_mode;

/* WindowFunction subset AggregateHash supports on the device (expression/window_function_expression.hpp). */
typedef enum hyb_aggregate_function {
  HYB_AGG_MIN = 0,
  HYB_AGG_MAX = 1,
  HYB_AGG_SUM = 2,
  HYB_AGG_AVG = 3,
  HYB_AGG_COUNT = 4,      /* COUNT(column): non-NULL rows */
  HYB_AGG_COUNT_STAR = 5, /* COUNT(*): INVALID_COLUMN_ID argument in the reference */
  HYB_AGG_COUNT_DISTINCT = 6,     /* unsupported -> CPU */
  HYB_AGG_STDDEV_SAMP = 7,        /* unsupported -> CPU */
  HYB_AGG_ANY = 8                 /* pseudo aggregate: handled like a group-by column by the shim */
} hyb_aggregate_function;

/* A typed scalar, already losslessly cast to the column type by the shim (utils/lossless_predicate_cast.hpp:20-60). */
typedef union hyb_value {
  int32_t i32;
  int64_t i64;
  float f32;
  double f64;
} hyb_value;

/* ------------------------------------------------------------------------------------------------------------------
 * Host-side segment / table views (borrowed pointers into the reference's pmr_vectors)
 * ---------------------------------------------------------------------------------------------------------------- */

/*
 * One segment of one chunk. Field use per encoding:
 *   UNENCODED          values = row_count x T; nulls optional.                         value_segment.hpp:84-85
 *   DICTIONARY         values = dictionary (dictionary_size x T, sorted, unique; NULL pointer for STRING),
 *                      attribute_vector = row_count value-IDs in vector_type layout,
 *                      NULL is encoded as value-ID == dictionary_size.                 dictionary_segment.hpp:88-90
 *   FRAME_OF_REFERENCE values = block minima (ceil(row_count/2048) x int32),
 *                      attribute_vector = row_count offsets in vector_type layout, nulls optional.
 *                                                                                      frame_of_reference_segment.hpp:94-97
 * nulls: one byte per row (0/1). The reference stores pmr_vector<bool> (bit-packed); the shim expands it once.
 * dictionary_codes (optional, DICTIONARY only): dictionary_size x uint64 chunk-independent group-by codes for this
 *   segment's dictionary entries — required to group by a STRING column (the shim computes them once per dictionary
 *   with the reference's own key scheme, aggregate_hash.cpp:818-925; O(dictionary) host work, not O(rows)).
 */
typedef struct hyb_segment_desc {
  int32_t encoding;    /* hyb_encoding */
  int32_t data_type;   /* hyb_data_type */
  int32_t vector_type; /* hyb_vector_type (DICTIONARY / FRAME_OF_REFERENCE) */
  int32_t bit_width;   /* BITPACKED only: bits per entry (1..32) */
  uint32_t row_count;
  uint32_t dictionary_size;
  const void* values;
  const uint8_t* nulls;
  const void* attribute_vector;
  const uint64_t* dictionary_codes;
} hyb_segment_desc;

/* A table = chunk_count x column_count segments, row-major by chunk: segments[chunk * column_count + column]. */
typedef struct hyb_table_view {
  uint32_t chunk_count;
  uint32_t column_count;
  const hyb_segment_desc* segments;
} hyb_table_view;

/* ------------------------------------------------------------------------------------------------------------------
 * Context and device column pool
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct hyb_context hyb_context; /* one per (process, device); owns streams, scratch and the column pool */
typedef uint64_t hyb_table_t;           /* device-resident table (handle into the pool) */
typedef uint64_t hyb_pos_list_t;        /* device-resident RowIDPosList set (one list per input chunk) */
typedef uint64_t hyb_join_result_t;
typedef uint64_t hyb_aggregate_result_t;

int hyb_abi_version(void);
const char* hyb_last_error(void);

/* device_index: CUDA ordinal (one process per GPU: pass LOCAL_RANK). */
int hyb_context_create(int device_index, hyb_context** out_context);
int hyb_context_destroy(hyb_context* context);
int hyb_device_count(int* out_count);
/* Block until all work queued by this context has finished. */
int hyb_context_synchronize(hyb_context* context);

/* Tuning / test knobs (not needed for correctness). The HYB_* environment variables of the same meaning are read once,
 * when the context is created; this call changes a knob afterwards. name/value pairs:
 *   "join_table"       = "auto" | "hash" | "direct" | "rank"   hash-table kind of hyb_join_hash (parity tests force each)
 *   "join_span"        = "0" | "1"   8192-row span kernels with shared-memory staged, coalesced PosList stores
 *   "join_rank"        = "ballot" | "match"
 *   "scan_bulk"        = "0" | "1"   cp.async.bulk + mbarrier input pipeline of the scan kernel
 *   "aggregate_stream" = "0" | "1"   bulk-staged streaming kernel for low-cardinality group-bys
 *   "aggregate_split"  = "0" | "1"   split a dictionary that exceeds shared memory over a CTA pair
 * Unknown names / values return HYB_ERR_INVALID. */
int hyb_context_set_option(hyb_context* context, const char* name, const char* value);

/* Pinned host memory for upload sources / result destinations (a pinned MemoryResource in the reference's terms,
 * cf. src/lib/memory/default_memory_resource.cpp:29-35). */
int hyb_host_alloc(size_t bytes, void** out_ptr);
int hyb_host_free(void* ptr);

/* Upload a whole table into the pool (H2D of every buffer the descriptors point to). */
int hyb_table_upload(hyb_context* context, const hyb_table_view* view, hyb_table_t* out_table);
/* Incremental variant: create an empty table, then append chunks (column_count segments each). */
int hyb_table_create(hyb_context* context, uint32_t column_count, hyb_table_t* out_table);
int hyb_table_append_chunk(hyb_context* context, hyb_table_t table, const hyb_segment_desc* segments);
/* Like hyb_table_append_chunk, but every pointer in `segments` already is a DEVICE pointer on this context's GPU
 * (e.g. tuples received through the multi-GPU exchange). Nothing is copied; the buffers are borrowed until the table is
 * dropped and must be 16-byte aligned with 64 readable bytes after their last element. */
int hyb_table_append_chunk_device(hyb_context* context, hyb_table_t table, const hyb_segment_desc* segments);
int hyb_table_drop(hyb_context* context, hyb_table_t table);

/*
 * Loading path: Hyrise's binary table format (BinaryParser::parse, import_export/binary/binary_parser.cpp:40-344 — what
 * hyriseBenchmarkTPCH caches under tpch_cached_tables/) -> host segments in this header's layout. Value / Dictionary /
 * FrameOfReference segments keep the reference's encoding (FixedWidthInteger and BitPacking vectors as stored), RunLength
 * segments are expanded, FixedStringDictionary and unencoded string segments become string dictionaries; every buffer sits
 * in a 256-byte aligned slot of a few large host blocks — pinned (`pinned` != 0 and a CUDA device is usable) so that
 * hyb_table_upload_binary moves the table with one DMA per block. HYB_ERR_NOT_FOUND: no such file; HYB_ERR_INVALID: what
 * the reference reports with Fail() (invalid encoding / vector type, truncated file); HYB_ERR_UNSUPPORTED: LZ4 segments.
 * The parse itself needs no GPU.
 */
typedef struct hyb_binary_table hyb_binary_table;
int hyb_binary_table_open(const char* path, int32_t pinned, hyb_binary_table** out_table);
void hyb_binary_table_close(hyb_binary_table* table);
int hyb_binary_table_info(const hyb_binary_table* table, uint32_t* out_chunk_size, uint32_t* out_chunk_count, uint32_t* out_column_count);
int hyb_binary_table_column(const hyb_binary_table* table, uint32_t column, const char** out_name, int32_t* out_data_type,
                            int32_t* out_nullable);
/* Segment descriptors (chunk-major) pointing into the table's host blocks; valid until hyb_binary_table_close. */
int hyb_binary_table_view(const hyb_binary_table* table, hyb_table_view* out_view);
/* Per chunk: the sorted-by information the file carries ({column id, SortMode} pairs, chunk.hpp). */
int hyb_binary_table_sorted_columns(const hyb_binary_table* table, uint32_t chunk, uint16_t* out_column_ids, uint8_t* out_sort_modes,
                                    uint32_t* out_count);
/* Host-resident dictionary of a string segment: entry i = chars[offsets[i] .. offsets[i + 1]). */
int hyb_binary_table_string_dictionary(const hyb_binary_table* table, uint32_t chunk, uint32_t column, const char** out_chars,
                                       const uint64_t** out_offsets, uint32_t* out_count);
/* hyb_scan_predicate.value_id_bounds for a string column of this table: per chunk DictionarySegment::lower_bound /
 * upper_bound of `value` (and of `value2` for BETWEEN; pass NULL otherwise): chunk_count x (2 | 4) entries. */
int hyb_binary_table_value_id_bounds(const hyb_binary_table* table, uint32_t column, const char* value, uint64_t value_length,
                                     const char* value2, uint64_t value2_length, uint32_t* out_bounds);

/*
 * Arena upload: when the segment buffers of one or more tables live inside a few large host blocks (a pinned
 * MemoryResource arena, cf. src/lib/memory/default_memory_resource.cpp:29-35 and Chunk::migrate, storage/chunk.hpp:119),
 * DMA each block ONCE and let tables point into the device copies instead of issuing one copy per segment buffer
 * (tens of thousands at SF 10). Buffers must be >= 16-byte aligned inside their block and followed by 64 readable bytes.
 * A buffer of the view that is not inside any block is copied individually, as hyb_table_upload does.
 */
typedef struct hyb_host_block {
  const void* base;
  uint64_t bytes;
} hyb_host_block;
typedef uint64_t hyb_block_set_t;
int hyb_blocks_upload(hyb_context* context, const hyb_host_block* blocks, uint32_t block_count,
                      hyb_block_set_t* out_block_set);
int hyb_table_upload_from_blocks(hyb_context* context, const hyb_table_view* view, hyb_block_set_t block_set,
                                 hyb_table_t* out_table);
/* Releases the device copies once no table references them any more. */
int hyb_blocks_free(hyb_context* context, hyb_block_set_t block_set);
/* The host blocks of a parsed binary table (pointer to an array owned by the table) and the one-call upload. */
int hyb_binary_table_blocks(const hyb_binary_table* table, hyb_host_block* out_blocks, uint32_t* out_count);
int hyb_table_upload_binary(hyb_context* context, const hyb_binary_table* table, hyb_table_t* out_table);
int hyb_table_info(hyb_context* context, hyb_table_t table, uint32_t* out_chunk_count, uint32_t* out_column_count,
                   uint64_t* out_row_count, uint64_t* out_device_bytes);

/* ------------------------------------------------------------------------------------------------------------------
 * TableScan
 * ---------------------------------------------------------------------------------------------------------------- */

/*
 * column <condition> value / column BETWEEN lower AND upper / column IS [NOT] NULL.
 * For STRING dictionary columns the device never sees strings: the shim passes, per chunk, the value-ID bounds the
 * reference computes with DictionarySegment::lower_bound/upper_bound (dictionary_segment.cpp:94-119):
 *   value_id_bounds[2*chunk+0] = lower_bound(lower), value_id_bounds[2*chunk+1] = upper_bound(lower)        (binary)
 *   value_id_bounds[4*chunk+0..3] = lower_bound(lower), upper_bound(lower), lower_bound(upper), upper_bound(upper)
 *                                                                                                           (between)
 * with HYB_INVALID_VALUE_ID for "past the end". For numeric dictionaries leave it NULL: bounds are computed on the
 * device from `lower`/`upper`.
 */
typedef struct hyb_scan_predicate {
  uint32_t column_id;
  int32_t condition; /* hyb_predicate_condition */
  hyb_value lower;   /* the value for binary conditions; lower bound for BETWEEN */
  hyb_value upper;   /* upper bound for BETWEEN */
  const uint32_t* value_id_bounds;
} hyb_scan_predicate;

/*
 * Predicate normalisation (host only): what TableScan::create_impl does to a literal before a scan implementation sees
 * it (table_scan.cpp:340-366, :399-441) with lossless_predicate_variant_cast (utils/lossless_predicate_cast.hpp:20-66,
 * .cpp:14-73), lossless_cast (lossless_cast.hpp:31-176) and flip_predicate_condition / between_to_conditions /
 * conditions_to_between (types.cpp:51-153). A shim whose optimizer hands it a literal of another type than the column
 * calls these and passes the results on in hyb_scan_predicate; HYB_ERR_UNSUPPORTED = "no lossless form": the reference
 * falls back to the ExpressionEvaluator, i.e. the CPU operator runs.
 *   `float_col < 3.1 (double)`  ->  `float_col <= 3.0999999f`        `int_col = 16.25`  ->  HYB_ERR_UNSUPPORTED
 * value_on_left: the predicate reads `literal <condition> column`; the returned condition is for `column <cond'> value`.
 */
typedef struct hyb_literal {
  int32_t data_type; /* hyb_data_type, numeric */
  hyb_value value;
} hyb_literal;
int hyb_flip_predicate_condition(int32_t condition, int32_t* out_condition);
int hyb_next_float_towards(double value, double towards, float* out_value, int32_t* out_possible);
int hyb_lossless_predicate_cast(int32_t condition, const hyb_literal* literal, int32_t column_type, int32_t value_on_left,
                                int32_t* out_condition, hyb_value* out_value);
int hyb_lossless_between_cast(int32_t condition, const hyb_literal* lower, const hyb_literal* upper, int32_t column_type,
                              int32_t* out_condition, hyb_value* out_lower, hyb_value* out_upper);

/*
 * Scan `table` and produce, per input chunk, the ascending list of matching RowIDs — what
 * AbstractTableScanImpl::scan_chunk returns for every chunk (table_scan.cpp:131). `input_filter` (0 = none) restricts
 * the scan to the positions of a previous scan's result on the same table (reference-table input with single-chunk
 * pos lists, abstract_dereferenced_column_table_scan_impl.cpp:38-46); the output then holds the referenced RowIDs
 * (table_scan.cpp:150-197).
 */
int hyb_table_scan(hyb_context* context, hyb_table_t table, const hyb_scan_predicate* predicate,
                   hyb_pos_list_t input_filter, hyb_pos_list_t* out_pos_list);

/* Total matches and per-chunk boundaries: out_chunk_offsets has chunk_count + 1 entries; chunk c's RowIDs are
 * [out_chunk_offsets[c], out_chunk_offsets[c+1]) of the flat list. Either pointer may be NULL. */
int hyb_pos_list_info(hyb_context* context, hyb_pos_list_t pos_list, uint64_t* out_total, uint32_t* out_chunk_count);
int hyb_pos_list_chunk_offsets(hyb_context* context, hyb_pos_list_t pos_list, uint64_t* out_chunk_offsets);
/* Copy RowIDs [begin, begin+count) of the flat list to host memory. */
int hyb_pos_list_copy(hyb_context* context, hyb_pos_list_t pos_list, uint64_t begin, uint64_t count,
                      hyb_row_id* out_row_ids);
int hyb_pos_list_free(hyb_context* context, hyb_pos_list_t pos_list);

/* ------------------------------------------------------------------------------------------------------------------
 * JoinHash (equi-join on one int32/int64 key column per side)
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct hyb_join_side {
  hyb_table_t table;
  uint32_t column_id;
  hyb_pos_list_t filter; /* 0 = all rows; else join only these positions (reference-table input) */
} hyb_join_side;

/*
 * build/probe sides are chosen by the caller exactly as JoinHash::_on_execute does (join_hash.cpp:139-155).
 * radix_bits: the reference's partition count (calculate_radix_bits, join_hash.cpp:70-114; pass -1 to have it computed
 * from the input sizes). It only fixes the ORDER and SLICING of the output: pairs are emitted grouped by
 * hash(key) & (2^radix_bits - 1) (std::hash<int> = identity), inside a partition in probe-row order, for one probe row
 * in build-row order — identical to probe() over radix-partitioned inputs (join_hash_steps.hpp:624-792).
 * Semi/Anti modes emit probe RowIDs only (probe_semi_anti, :794-922).
 */
int hyb_join_hash(hyb_context* context, const hyb_join_side* build, const hyb_join_side* probe, int32_t mode,
                  int32_t radix_bits, hyb_join_result_t* out_result);

/* out_pair_count: number of emitted rows; out_partition_count: 2^radix_bits. */
int hyb_join_result_info(hyb_context* context, hyb_join_result_t result, uint64_t* out_pair_count,
                         uint32_t* out_partition_count, int32_t* out_radix_bits);
/* partition_count + 1 offsets into the flat pair list (output of partition p = [off[p], off[p+1])). */
int hyb_join_result_partition_offsets(hyb_context* context, hyb_join_result_t result, uint64_t* out_offsets);
/* Copy pairs [begin, begin+count). out_build_row_ids may be NULL (always for Semi/Anti). */
int hyb_join_result_copy(hyb_context* context, hyb_join_result_t result, uint64_t begin, uint64_t count,
                         hyb_row_id* out_build_row_ids, hyb_row_id* out_probe_row_ids);
int hyb_join_result_free(hyb_context* context, hyb_join_result_t result);

/*
 * Composing operators on the device: one side of a join result as a PosList on that side's table — the
 * reference table a following operator of the plan consumes (join_output_writing.cpp:95-200: output ReferenceSegments always
 * point at the ORIGINAL data table, which is what the RowIDs of a result already do). side: 0 = build, 1 = probe. The list
 * is in result order (not table order: `hyb_pos_list_info` reports one chunk), may hold NULL_ROW_ID for outer joins, and is
 * accepted as `filter` / `input_filter` by hyb_table_scan, hyb_join_hash and hyb_aggregate_hash like a scan's PosList:
 * scans keep the input order (abstract_dereferenced_column_table_scan_impl.cpp:49-86: matches are positions of the input
 * list, mapped back to the referenced RowIDs by table_scan.cpp:150-197), NULL rows never match / never join.
 */
int hyb_join_result_pos_list(hyb_context* context, hyb_join_result_t result, int32_t side, hyb_pos_list_t* out_pos_list);

/*
 * The chunks write_output_chunks would cut the result into (join_output_writing.cpp:205-340): the PosLists probe() emits —
 * one per non-empty radix partition, cut every PROBE_SIZE_PER_CHUNK = 131 070 probe elements of the partition
 * (join_hash_steps.hpp:47, :655) — merged while they hold fewer than MIN_SIZE = 1000 rows and the merge stays below
 * MAX_SIZE = 4000 (:255-296). out_offsets receives *inout_count + 1 offsets into the flat pair list; call with
 * out_offsets = NULL to get the count. (Slices are cut over the emitted rows of a partition: without secondary
 * predicates and Bloom-filter false positives that is the reference's element count for PK-FK joins; see DESIGN.md.)
 */
int hyb_join_result_output_chunks(hyb_context* context, hyb_join_result_t result, uint64_t* out_offsets, uint32_t* inout_count);

/*
 * Multi-GPU radix exchange, step 1 (materialize_input, join_hash_steps.hpp:274-420, without the Bloom filter): write one
 * join side as {key, RowID} tuples into caller-provided DEVICE buffers of hyb_join_side_positions() elements each:
 * out_keys[i] = key as int64, out_row_ids[i] = chunk_id + chunk_id_base | chunk_offset << 32, in row order. Rows with a
 * NULL key get out_row_ids[i] = -1 (Inner/Semi joins drop them before the exchange). The host layer partitions the
 * tuples by hash(key) & (world - 1), exchanges them with one all-to-all (NCCL) and joins what it received with
 * hyb_table_append_chunk_device + hyb_join_hash.
 */
int hyb_join_side_positions(hyb_context* context, const hyb_join_side* side, uint64_t* out_positions);
int hyb_join_materialize(hyb_context* context, const hyb_join_side* side, uint32_t chunk_id_base, void* out_keys_device,
                         void* out_row_ids_device);

/*
 * Multi-GPU radix exchange, sender side in one step (materialize_input + partition_by_radix, join_hash_steps.hpp:274-420
 * and :422-560, with the RANK that owns a radix partition as the partition): writes the side's non-NULL {key, RowID}
 * tuples into caller-provided DEVICE buffers of hyb_join_side_positions() elements each, grouped by
 * key & (partition_count - 1) and in row order inside a group (a stable split, so that the receiving rank sees the rows
 * of every source rank in their original order). out_keys[i] = key as int64, out_row_ids[i] = {chunk_id + chunk_id_base,
 * chunk_offset}. out_partition_offsets (HOST, partition_count + 1 entries): group p occupies [offsets[p], offsets[p + 1]).
 * The host layer sends group p to rank p with one all-to-all (NCCL) and joins what it received with
 * hyb_table_append_chunk_device + hyb_join_hash. partition_count: power of two <= 256.
 */
int hyb_join_partition(hyb_context* context, const hyb_join_side* side, uint32_t partition_count, uint32_t chunk_id_base,
                       void* out_keys_device, void* out_row_ids_device, uint64_t* out_partition_offsets);

/*
 * The same split fused with the exchange: ONE ranked-write kernel stores every group straight into the memory of the rank
 * that owns it (NVLink peer-to-peer stores through CUDA IPC mappings) — no send buffer, no NCCL payload transfer.
 * Between the count pass and the write pass the library calls `exchange(user, counts, dest_keys, dest_row_ids)` with this
 * rank's tuple count per destination rank; the host layer all-gathers the counts (the only collective of the exchange,
 * world_size^2 integers) and answers with, for every destination rank d, the DEVICE addresses (valid in THIS process:
 * hyb_exchange_arena_open) where this rank's group d starts inside rank d's receive arena. The call returns after this
 * rank's stores have completed; a barrier across ranks then makes every arena complete. world_size: power of two <= 16.
 */
#define HYB_IPC_HANDLE_BYTES 64
typedef int (*hyb_exchange_fn)(void* user, const uint64_t* counts, void** dest_keys, void** dest_row_ids);
int hyb_join_partition_push(hyb_context* context, const hyb_join_side* side, uint32_t world_size, uint32_t chunk_id_base,
                            hyb_exchange_fn exchange, void* user);
/* Receive arenas: plain device memory of this rank exported to / imported from the other ranks of the node. */
int hyb_exchange_arena_create(hyb_context* context, uint64_t bytes, void** out_device_ptr, void* out_ipc_handle);
int hyb_exchange_arena_open(hyb_context* context, const void* ipc_handle, void** out_peer_ptr);
int hyb_exchange_arena_close(hyb_context* context, void* peer_ptr);
int hyb_exchange_arena_destroy(hyb_context* context, void* device_ptr);

/*
 * Peer groups: the multi-GPU operators without host orchestration. One process per GPU; every rank creates its group
 * (allocating its exchange arena: a control block plus four regions of tuple_capacity tuples), the host layer
 * all-gathers the IPC handles once (the only use of torch.distributed / NCCL besides verification), every rank connects.
 * After that a distributed operator is ONE C-ABI call per rank: counts, statistics and partial groups travel through the
 * arenas with NVLink peer stores issued by the library's kernels, ranks order themselves with epoch flags polled on the
 * device, and the only host round trip is the one that reads the world x world count matrix. Every rank must make the
 * same distributed calls in the same order. world: power of two <= 16.
 */
typedef uint64_t hyb_peer_group_t;
int hyb_peer_group_create(hyb_context* context, uint32_t rank, uint32_t world, uint64_t tuple_capacity, void* out_ipc_handle,
                          hyb_peer_group_t* out_group);
/* all_ipc_handles: world x HYB_IPC_HANDLE_BYTES, in rank order (the own entry is ignored). */
int hyb_peer_group_connect(hyb_context* context, hyb_peer_group_t group, const void* all_ipc_handles);
int hyb_peer_group_destroy(hyb_context* context, hyb_peer_group_t group);

/*
 * Inner JoinHash over ranks. The ranks first exchange the [min, max] of their build and probe keys through
 * the control blocks. Co-located shards (no rank's probe range touches another rank's build range) are joined locally and
 * the RowIDs are globalised by adding build_chunk_base / probe_chunk_base to the chunk ids: every rank then holds its slice
 * of EVERY radix partition, and partition p of the global result is the concatenation of the ranks' slices in rank order
 * (hyb_distributed_stats.colocated == 1). Otherwise both sides are split by the rank that owns the radix partition
 * (key & (world - 1), the reference's hash(key) & mask with the identity hash) and pushed straight into the owners' arenas
 * by the fused split + NVLink store kernel; every rank joins what it received and emits the GLOBAL RowIDs that travelled
 * with the keys: a rank ends up with the partitions p = rank (mod world) of the reference-ordered result, in the
 * reference's order (colocated == 0). HYB_JOIN_COLOCATED=0 / option "join_colocated" forces the exchange.
 */
int hyb_join_hash_distributed(hyb_context* context, hyb_peer_group_t group, const hyb_join_side* build,
                              const hyb_join_side* probe, uint32_t build_chunk_base, uint32_t probe_chunk_base,
                              int32_t radix_bits, hyb_join_result_t* out_result);

/*
 * AggregateHash over ranks: every rank pre-aggregates its shard (AVG travels as SUM and COUNT). The ranks agree on the form:
 * low cardinality (every rank's partial groups fit one 32 KB block): the blocks are stored into every peer's arena and
 * merged by all ranks in rank order — every rank returns the complete, bit-identical result; high cardinality (Q3 / Q18
 * shapes): the partial groups are partitioned by the hash of their key, pushed into the owners' tuple regions over NVLink
 * and merged by the owner — a rank returns the groups it owns (hyb_distributed_stats.aggregate_partitioned == 1).
 * Representative RowIDs refer to the global table (chunk ids shifted by chunk_id_base); position_base = number of rows of
 * the global table before this rank's shard (first-appearance order across ranks).
 */
int hyb_aggregate_hash_distributed(hyb_context* context, hyb_peer_group_t group, const struct hyb_aggregate_query* query,
                                   uint32_t chunk_id_base, uint64_t position_base, hyb_aggregate_result_t* out_result);

/* Phases of the last distributed call on this rank (CUDA-event times on the context stream). */
typedef struct hyb_distributed_stats {
  float split_count_ms;      /* per-destination counts of both sides + publishing them to the peers */
  float count_wait_ms;       /* waiting for every peer's counts (device-side flag wait) incl. the host read of the matrix */
  float push_ms;             /* fused split + NVLink stores of both sides */
  float done_wait_ms;        /* waiting for every peer's stores */
  float local_ms;            /* local join (or pre-aggregation) */
  float finish_ms;           /* RowID translation / merge */
  uint64_t tuples_sent;      /* to other ranks */
  uint64_t tuples_received;  /* from all ranks, own included */
  uint64_t nvlink_bytes;     /* bytes this rank stored into peer memory */
  uint32_t colocated;        /* join: 1 = the ranks' key ranges did not overlap across ranks, every rank joined its own shards
                                (result layout: every rank holds its slice of EVERY partition, global order = rank order
                                inside a partition); 0 = radix exchange (a partition lives on rank partition % world) */
  uint32_t aggregate_partitioned; /* aggregate: 1 = high-cardinality form — the partial groups were partitioned by the hash of
                                     their key and this rank's result holds the groups it owns (every group on exactly one
                                     rank, in the reference's order among the rank's groups; the global order is the merge
                                     of the ranks' results by representative RowID / key); 0 = every rank holds the complete
                                     result */
} hyb_distributed_stats;
int hyb_peer_group_stats(hyb_context* context, hyb_peer_group_t group, hyb_distributed_stats* out_stats);

/* ------------------------------------------------------------------------------------------------------------------
 * AggregateHash (with optionally fused scan predicates and Projection arithmetic)
 * ---------------------------------------------------------------------------------------------------------------- */

/* Arithmetic expression over columns in reverse Polish notation, evaluated with the reference's type rules
 * (expression/expression_utils.cpp:172-205: float∘float→float, int∘float→float, int∘int→int ...). A plain column is
 * the one-op program {HYB_EXPR_COLUMN}. */
typedef enum hyb_expr_op {
  HYB_EXPR_COLUMN = 0,   /* push column `column_id` */
  HYB_EXPR_LITERAL = 1,  /* push literal (type in literal_type, value in literal) */
  HYB_EXPR_ADD = 2,
  HYB_EXPR_SUB = 3,
  HYB_EXPR_MUL = 4,
  HYB_EXPR_DIV = 5
} hyb_expr_op;

typedef struct hyb_expr_node {
  int32_t op;           /* hyb_expr_op */
  uint32_t column_id;   /* COLUMN */
  int32_t literal_type; /* LITERAL: hyb_data_type */
  hyb_value literal;    /* LITERAL */
} hyb_expr_node;

#define HYB_MAX_EXPR_NODES 16
#define HYB_MAX_GROUPBY_COLUMNS 8
#define HYB_MAX_AGGREGATES 16
#define HYB_MAX_FUSED_PREDICATES 8

typedef struct hyb_aggregate_def {
  int32_t function; /* hyb_aggregate_function */
  uint32_t node_count; /* 0 for COUNT(*) */
  hyb_expr_node nodes[HYB_MAX_EXPR_NODES];
} hyb_aggregate_def;

typedef struct hyb_aggregate_query {
  hyb_table_t table;
  hyb_pos_list_t filter;           /* 0 = all rows; else aggregate only these positions */
  uint32_t predicate_count;        /* fused conjunctive ColumnVsValue/Between/IsNull predicates (may be 0) */
  const hyb_scan_predicate* predicates;
  uint32_t groupby_count;
  const uint32_t* groupby_column_ids;
  uint32_t aggregate_count;
  const hyb_aggregate_def* aggregates;
} hyb_aggregate_query;

/*
 * Groups are reported in the order AggregateHash emits them: first appearance in row order, or — when the reference's
 * "immediate key" shortcut applies (single int32 group-by column with a dense key range, aggregate_hash.cpp:781-804)
 * — ascending key order with the NULL group first. Representative RowIDs mirror aggregate_hash.cpp:367,394.
 */
int hyb_aggregate_hash(hyb_context* context, const hyb_aggregate_query* query, hyb_aggregate_result_t* out_result);

int hyb_aggregate_result_info(hyb_context* context, hyb_aggregate_result_t result, uint64_t* out_group_count,
                              int32_t* out_used_immediate_keys);
/* One representative RowID per group (what write_groupby_output turns into ReferenceSegments, :421-537). */
int hyb_aggregate_result_row_ids(hyb_context* context, hyb_aggregate_result_t result, hyb_row_id* out_row_ids);
/*
 * Values of aggregate `aggregate_index` for all groups. Result type follows WindowFunctionTraits
 * (aggregate/window_function_traits.hpp:14-77): COUNT* -> int64; SUM(int) -> int64; SUM(float/double) -> double;
 * AVG -> double; MIN/MAX -> the column type. `out_values` receives group_count elements of that type (8 bytes for
 * int64/double, 4 for int32/float); out_value_type tells which. out_nulls (optional) gets 1 where the reference writes
 * NULL (aggregate_count == 0).
 */
int hyb_aggregate_result_values(hyb_context* context, hyb_aggregate_result_t result, uint32_t aggregate_index,
                                void* out_values, uint8_t* out_nulls, int32_t* out_value_type);
int hyb_aggregate_result_free(hyb_context* context, hyb_aggregate_result_t result);

/*
 * Sort + Limit over an aggregate's output (operators/sort.cpp, operators/limit.cpp — `ORDER BY <aggregate> DESC LIMIT k`,
 * the tail of TPC-H Q3 / Q10 / Q18): the k groups with the largest (descending != 0) or smallest values of aggregate
 * `aggregate_index`, ties in group order (Sort is stable). The values are sorted on the device (top-k selection kernel);
 * out_group_indexes receives min(k, group_count) indexes into the result's groups. NULL values sort last.
 */
int hyb_aggregate_result_top_k(hyb_context* context, hyb_aggregate_result_t result, uint32_t aggregate_index, uint32_t k,
                               int32_t descending, uint32_t* out_group_indexes, uint32_t* out_count);

/* ------------------------------------------------------------------------------------------------------------------
 * Instrumentation (fills the reference's performance_data, operators/operator_performance_data.hpp:46-97)
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct hyb_operator_stats {
  float device_ms;            /* CUDA-event time of the operator's kernels on the context stream */
  float dominant_kernel_ms;   /* the bandwidth-bound kernel alone (scan / probe / aggregate kernel) */
  uint32_t kernel_launches;   /* kernels launched by the last operator call */
  uint32_t reserved;
  uint64_t algorithmic_bytes; /* compulsory traffic of the dominant kernel */
  uint64_t input_rows;
  uint64_t output_rows;
} hyb_operator_stats;

/* Stats of the last operator call made on this context by the calling thread. */
int hyb_last_operator_stats(hyb_context* context, hyb_operator_stats* out_stats);

/* Multi-GPU support: raw device pointers of result buffers so the host layer can hand them to NCCL
 * (torch.distributed) for the radix exchange. Pointers stay valid until the result is freed. */
int hyb_pos_list_device_ptr(hyb_context* context, hyb_pos_list_t pos_list, void** out_device_row_ids);
int hyb_join_result_device_ptrs(hyb_context* context, hyb_join_result_t result, void** out_build_row_ids,
                                void** out_probe_row_ids);
/* The context's CUDA stream (cudaStream_t as void*), so torch can order its collectives after our kernels. */
int hyb_context_stream(hyb_context* context, void** out_stream);

#ifdef __cplusplus
}
#endif

#endif /* HYRISE_B200_H */
