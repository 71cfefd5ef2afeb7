// CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle.h). A restatement of the reference algorithms; every function cites
// the reference file:line it follows (paths relative to /root/reference/src/lib). Not linked into the product.
#include "oracle.h"

#include <algorithm>
#include <optional>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(int status, const std::string& message) {
  g_error = message;
  return status;
}

#define ORC_CHECK(cond, status, message) \
  do {                                   \
    if (!(cond)) return fail((status), (message)); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// Scheduler stand-in: Hyrise spawns one JobTask per chunk / partition and waits (scheduler/abstract_scheduler.cpp:53-63).
// ---------------------------------------------------------------------------------------------------------------------
void run_jobs(size_t job_count, int threads, const std::function<void(size_t)>& job) {
  if (threads <= 1 || job_count <= 1) {
    for (size_t index = 0; index < job_count; ++index) job(index);
    return;
  }
  std::atomic<size_t> next{0};
  const auto worker = [&]() {
    while (true) {
      const size_t index = next.fetch_add(1);
      if (index >= job_count) return;
      job(index);
    }
  };
  std::vector<std::thread> pool;
  const size_t count = std::min<size_t>(threads, job_count);
  pool.reserve(count);
  for (size_t t = 0; t + 1 < count; ++t) pool.emplace_back(worker);
  worker();
  for (auto& thread : pool) thread.join();
}

// ---------------------------------------------------------------------------------------------------------------------
// Compressed vectors
// ---------------------------------------------------------------------------------------------------------------------
// FixedWidthIntegerVector<T>::get / BitPackingVector (compact_iterator.hpp:218-235: value i lives at bit i*b of a uint64
// word stream, least significant bit first, continuing in the next word).
inline uint32_t vector_get(const void* data, int32_t vector_type, int32_t bits, uint32_t index) {
  switch (vector_type) {
    case HYB_VEC_FIXED_1B:
      return static_cast<const uint8_t*>(data)[index];
    case HYB_VEC_FIXED_2B:
      return static_cast<const uint16_t*>(data)[index];
    case HYB_VEC_FIXED_4B:
      return static_cast<const uint32_t*>(data)[index];
    default: {
      const auto* words = static_cast<const uint64_t*>(data);
      const uint64_t bit = static_cast<uint64_t>(index) * static_cast<uint64_t>(bits);
      const uint64_t word = bit / 64;
      const uint32_t shift = static_cast<uint32_t>(bit % 64);
      const uint64_t mask = bits >= 64 ? ~uint64_t{0} : ((uint64_t{1} << bits) - 1);
      uint64_t value = words[word] >> shift;
      if (shift + bits > 64) value |= words[word + 1] << (64 - shift);
      return static_cast<uint32_t>(value & mask);
    }
  }
}

// Segment accessors -------------------------------------------------------------------------------------------------
inline bool segment_is_null(const hyb_segment_desc& segment, uint32_t offset) {
  if (segment.encoding == HYB_ENC_DICTIONARY) {
    // NULL is value-ID == dictionary size (dictionary_segment.cpp:139-141).
    return vector_get(segment.attribute_vector, segment.vector_type, segment.bit_width, offset) ==
           segment.dictionary_size;
  }
  return segment.nulls && segment.nulls[offset];
}

// Value at `offset`. For NULL positions the reference iterators yield: the stored value for ValueSegments
// (value_segment_iterable.hpp; load_table stores T{} for NULL), T{} for dictionaries
// (dictionary_segment_iterable.hpp:116), minimum + offset for FrameOfReference
// (frame_of_reference_segment_iterable.hpp:131-139). JoinHash hashes that value for NULL probe rows it keeps.
template <typename T>
inline T segment_value(const hyb_segment_desc& segment, uint32_t offset) {
  switch (segment.encoding) {
    case HYB_ENC_UNENCODED:
      return static_cast<const T*>(segment.values)[offset];
    case HYB_ENC_DICTIONARY: {
      const uint32_t value_id = vector_get(segment.attribute_vector, segment.vector_type, segment.bit_width, offset);
      if (value_id >= segment.dictionary_size) return T{};
      return static_cast<const T*>(segment.values)[value_id];
    }
    default: {  // frame_of_reference_segment.hpp:65-73
      if constexpr (std::is_same_v<T, int32_t>) {
        const int32_t minimum = static_cast<const int32_t*>(segment.values)[offset / HYB_FOR_BLOCK_SIZE];
        return static_cast<int32_t>(static_cast<uint32_t>(minimum) +
                                    vector_get(segment.attribute_vector, segment.vector_type, segment.bit_width, offset));
      } else {
        return T{};
      }
    }
  }
}

template <typename Functor>
auto with_type(int32_t data_type, const Functor& functor) {
  switch (data_type) {
    case HYB_TYPE_INT32:
      return functor(int32_t{});
    case HYB_TYPE_INT64:
      return functor(int64_t{});
    case HYB_TYPE_FLOAT32:
      return functor(float{});
    default:
      return functor(double{});
  }
}

template <typename T>
T value_of(const hyb_value& value) {
  if constexpr (std::is_same_v<T, int32_t>) return value.i32;
  if constexpr (std::is_same_v<T, int64_t>) return value.i64;
  if constexpr (std::is_same_v<T, float>) return value.f32;
  if constexpr (std::is_same_v<T, double>) return value.f64;
}

inline const hyb_segment_desc& segment_at(const hyb_table_view* table, uint32_t chunk, uint32_t column) {
  return table->segments[static_cast<size_t>(chunk) * table->column_count + column];
}

bool is_between_condition(int32_t c) { return c >= HYB_PRED_BETWEEN_INCLUSIVE && c <= HYB_PRED_BETWEEN_EXCLUSIVE; }
bool is_lower_inclusive_between(int32_t c) {  // types.cpp
  return c == HYB_PRED_BETWEEN_INCLUSIVE || c == HYB_PRED_BETWEEN_UPPER_EXCLUSIVE;
}
bool is_upper_inclusive_between(int32_t c) {
  return c == HYB_PRED_BETWEEN_INCLUSIVE || c == HYB_PRED_BETWEEN_LOWER_EXCLUSIVE;
}

using RowIDs = std::vector<hyb_row_id>;
constexpr uint32_t INVALID_VALUE_ID = HYB_INVALID_VALUE_ID;

// ---------------------------------------------------------------------------------------------------------------------
// TableScan
// ---------------------------------------------------------------------------------------------------------------------

// DictionarySegment::lower_bound / upper_bound (dictionary_segment.cpp:94-119): INVALID_VALUE_ID when past the end.
template <typename T>
uint32_t dictionary_lower_bound(const hyb_segment_desc& segment, T value) {
  const T* begin = static_cast<const T*>(segment.values);
  const T* end = begin + segment.dictionary_size;
  const T* it = std::lower_bound(begin, end, value);
  return it == end ? INVALID_VALUE_ID : static_cast<uint32_t>(it - begin);
}
template <typename T>
uint32_t dictionary_upper_bound(const hyb_segment_desc& segment, T value) {
  const T* begin = static_cast<const T*>(segment.values);
  const T* end = begin + segment.dictionary_size;
  const T* it = std::upper_bound(begin, end, value);
  return it == end ? INVALID_VALUE_ID : static_cast<uint32_t>(it - begin);
}

struct Bounds {
  uint32_t lower_bound;  // lower_bound(value)
  uint32_t upper_bound;  // upper_bound(value)
};

Bounds bounds_for(const hyb_segment_desc& segment, const hyb_value& value, const uint32_t* host_bounds) {
  if (host_bounds) return {host_bounds[0], host_bounds[1]};  // string dictionaries: computed by the caller
  return with_type(segment.data_type, [&](auto tag) {
    using T = decltype(tag);
    return Bounds{dictionary_lower_bound<T>(segment, value_of<T>(value)),
                  dictionary_upper_bound<T>(segment, value_of<T>(value))};
  });
}

// AbstractTableScanImpl::_scan_with_iterators (abstract_table_scan_impl.hpp:56-84), scalar form: iterate positions
// (all rows, or the rows of `position_filter`), emit RowID{chunk_id, position index} for matches.
template <bool CheckForNull, typename Functor>
void scan_with_iterators(const hyb_segment_desc& segment, uint32_t chunk_id, const RowIDs* position_filter,
                         const Functor& functor, RowIDs& matches) {
  const uint32_t count = position_filter ? static_cast<uint32_t>(position_filter->size()) : segment.row_count;
  for (uint32_t position = 0; position < count; ++position) {
    const uint32_t offset = position_filter ? (*position_filter)[position].chunk_offset : position;
    if ((!CheckForNull || !segment_is_null(segment, offset)) && functor(offset)) {
      matches.push_back(hyb_row_id{chunk_id, position});
    }
  }
}

void add_all(uint32_t chunk_id, uint32_t count, RowIDs& matches) {
  for (uint32_t offset = 0; offset < count; ++offset) matches.push_back(hyb_row_id{chunk_id, offset});
}

bool segment_may_contain_nulls(const hyb_segment_desc& segment) {
  // Stand-in for `_column_is_nullable` (table column definition): a dictionary segment of a nullable column may hold the
  // NULL value-ID. Treating every dictionary column as nullable only disables a fast path, results are identical.
  return segment.encoding == HYB_ENC_DICTIONARY || segment.nulls != nullptr;
}

// ColumnVsValueTableScanImpl::_scan_dictionary_segment (column_vs_value_table_scan_impl.cpp:89-180)
void scan_dictionary_vs_value(const hyb_segment_desc& segment, uint32_t chunk_id, int32_t condition,
                              const hyb_value& value, const uint32_t* host_bounds, const RowIDs* position_filter,
                              RowIDs& matches) {
  const Bounds bounds = bounds_for(segment, value, host_bounds);
  // _get_search_value_id (:206-226)
  uint32_t search_value_id;
  switch (condition) {
    case HYB_PRED_EQUALS:
    case HYB_PRED_NOT_EQUALS:
    case HYB_PRED_LESS_THAN:
    case HYB_PRED_GREATER_THAN_EQUALS:
      search_value_id = bounds.lower_bound;
      break;
    default:
      search_value_id = bounds.upper_bound;
      break;
  }
  // value_of_value_id(search_value_id) == value  <=>  lower_bound != upper_bound (the value is in the dictionary)
  const bool value_in_dictionary = bounds.lower_bound != bounds.upper_bound;
  const uint32_t unique_values_count = segment.dictionary_size;

  // _value_matches_all (:228-250)
  bool matches_all = false, matches_none = false;
  switch (condition) {
    case HYB_PRED_EQUALS:
      matches_all = search_value_id != INVALID_VALUE_ID && value_in_dictionary && unique_values_count == 1;
      matches_none = search_value_id == INVALID_VALUE_ID || !value_in_dictionary;
      break;
    case HYB_PRED_NOT_EQUALS:
      matches_all = search_value_id == INVALID_VALUE_ID || !value_in_dictionary;
      matches_none = search_value_id != INVALID_VALUE_ID && value_in_dictionary && unique_values_count == 1;
      break;
    case HYB_PRED_LESS_THAN:
    case HYB_PRED_LESS_THAN_EQUALS:
      matches_all = search_value_id == INVALID_VALUE_ID;
      matches_none = search_value_id == 0;
      break;
    default:
      matches_all = search_value_id == 0;
      matches_none = search_value_id == INVALID_VALUE_ID;
      break;
  }
  const auto value_id_at = [&](uint32_t offset) {
    return vector_get(segment.attribute_vector, segment.vector_type, segment.bit_width, offset);
  };
  if (matches_all) {
    if (segment_may_contain_nulls(segment)) {
      scan_with_iterators<true>(segment, chunk_id, position_filter, [](uint32_t) { return true; }, matches);
    } else {
      add_all(chunk_id, position_filter ? static_cast<uint32_t>(position_filter->size()) : segment.row_count, matches);
    }
    return;
  }
  if (matches_none) return;

  // _with_operator_for_dict_segment_scan (column_vs_value_table_scan_impl.hpp:58-81) + NULL handling (:162-179)
  switch (condition) {
    case HYB_PRED_EQUALS:
      scan_with_iterators<false>(segment, chunk_id, position_filter,
                                 [&](uint32_t o) { return value_id_at(o) == search_value_id; }, matches);
      break;
    case HYB_PRED_NOT_EQUALS:
      scan_with_iterators<true>(segment, chunk_id, position_filter,
                                [&](uint32_t o) { return value_id_at(o) != search_value_id; }, matches);
      break;
    case HYB_PRED_LESS_THAN:
    case HYB_PRED_LESS_THAN_EQUALS:
      scan_with_iterators<false>(segment, chunk_id, position_filter,
                                 [&](uint32_t o) { return value_id_at(o) < search_value_id; }, matches);
      break;
    default:
      scan_with_iterators<true>(segment, chunk_id, position_filter,
                                [&](uint32_t o) { return value_id_at(o) >= search_value_id; }, matches);
      break;
  }
}

// ColumnVsValueTableScanImpl::_scan_generic_segment (:64-87) with with_comparator (type_comparison.hpp:160-180)
void scan_generic_vs_value(const hyb_segment_desc& segment, uint32_t chunk_id, int32_t condition, const hyb_value& value,
                           const RowIDs* position_filter, RowIDs& matches) {
  with_type(segment.data_type, [&](auto tag) {
    using T = decltype(tag);
    const T typed_value = value_of<T>(value);
    const auto scan = [&](auto comparator) {
      scan_with_iterators<true>(segment, chunk_id, position_filter,
                                [&](uint32_t o) { return comparator(segment_value<T>(segment, o), typed_value); },
                                matches);
    };
    switch (condition) {
      case HYB_PRED_EQUALS:
        scan(std::equal_to<T>{});
        break;
      case HYB_PRED_NOT_EQUALS:
        scan(std::not_equal_to<T>{});
        break;
      case HYB_PRED_LESS_THAN:
        scan(std::less<T>{});
        break;
      case HYB_PRED_LESS_THAN_EQUALS:
        scan(std::less_equal<T>{});
        break;
      case HYB_PRED_GREATER_THAN:
        scan(std::greater<T>{});
        break;
      default:
        scan(std::greater_equal<T>{});
        break;
    }
    return 0;
  });
}

// ColumnBetweenTableScanImpl::_scan_dictionary_segment (column_between_table_scan_impl.cpp:112-194)
void scan_dictionary_between(const hyb_segment_desc& segment, uint32_t chunk_id, int32_t condition,
                             const hyb_value& lower, const hyb_value& upper, const uint32_t* host_bounds,
                             const RowIDs* position_filter, RowIDs& matches) {
  const Bounds lower_bounds = bounds_for(segment, lower, host_bounds);
  const Bounds upper_bounds = bounds_for(segment, upper, host_bounds ? host_bounds + 2 : nullptr);
  uint32_t lower_bound_value_id =
      is_lower_inclusive_between(condition) ? lower_bounds.lower_bound : lower_bounds.upper_bound;
  uint32_t upper_bound_value_id =
      is_upper_inclusive_between(condition) ? upper_bounds.upper_bound : upper_bounds.lower_bound;

  if (lower_bound_value_id == 0 && upper_bound_value_id == INVALID_VALUE_ID) {
    if (segment_may_contain_nulls(segment)) {
      scan_with_iterators<true>(segment, chunk_id, position_filter, [](uint32_t) { return true; }, matches);
    } else {
      add_all(chunk_id, position_filter ? static_cast<uint32_t>(position_filter->size()) : segment.row_count, matches);
    }
    return;
  }
  if (lower_bound_value_id == INVALID_VALUE_ID || lower_bound_value_id >= upper_bound_value_id) return;
  if (upper_bound_value_id == INVALID_VALUE_ID) upper_bound_value_id = segment.dictionary_size;

  // with_between_comparator(BetweenUpperExclusive, ...) on integral value ids (type_comparison.hpp:115-133)
  const uint32_t lower_bound = lower_bound_value_id;
  const uint32_t value_difference = (upper_bound_value_id - 1) - lower_bound;
  scan_with_iterators<false>(
      segment, chunk_id, position_filter,
      [&](uint32_t o) {
        const uint32_t value_id = vector_get(segment.attribute_vector, segment.vector_type, segment.bit_width, o);
        return static_cast<uint32_t>(value_id - lower_bound) <= value_difference;
      },
      matches);
}

// ColumnBetweenTableScanImpl::_scan_generic_segment (:71-110)
void scan_generic_between(const hyb_segment_desc& segment, uint32_t chunk_id, int32_t condition, const hyb_value& lower,
                          const hyb_value& upper, const RowIDs* position_filter, RowIDs& matches) {
  with_type(segment.data_type, [&](auto tag) {
    using T = decltype(tag);
    const T left = value_of<T>(lower);
    const T right = value_of<T>(upper);
    if constexpr (std::is_integral_v<T>) {
      // (:88-97) computed in a wider type here so INT_MIN/INT_MAX bounds cannot overflow
      const __int128 difference = static_cast<__int128>(right) - static_cast<__int128>(left) -
                                  !is_lower_inclusive_between(condition) - !is_upper_inclusive_between(condition);
      if (difference < 0) return 0;
    }
    const bool lower_inclusive = is_lower_inclusive_between(condition);
    const bool upper_inclusive = is_upper_inclusive_between(condition);
    scan_with_iterators<true>(
        segment, chunk_id, position_filter,
        [&](uint32_t o) {
          const T v = segment_value<T>(segment, o);
          const bool above = lower_inclusive ? v >= left : v > left;
          const bool below = upper_inclusive ? v <= right : v < right;
          return above && below;
        },
        matches);
    return 0;
  });
}

// ColumnIsNullTableScanImpl (column_is_null_table_scan_impl.cpp): invert ^ is_null
void scan_is_null(const hyb_segment_desc& segment, uint32_t chunk_id, int32_t condition, const RowIDs* position_filter,
                  RowIDs& matches) {
  const bool invert = condition == HYB_PRED_IS_NOT_NULL;
  scan_with_iterators<false>(segment, chunk_id, position_filter,
                             [&](uint32_t o) { return invert != segment_is_null(segment, o); }, matches);
}

// AbstractDereferencedColumnTableScanImpl::scan_chunk (abstract_dereferenced_column_table_scan_impl.cpp:19-46)
void scan_chunk(const hyb_table_view* table, uint32_t chunk_id, const hyb_scan_predicate* predicate,
                const RowIDs* position_filter, RowIDs& matches) {
  const auto& segment = segment_at(table, chunk_id, predicate->column_id);
  const int32_t condition = predicate->condition;
  const bool dictionary = segment.encoding == HYB_ENC_DICTIONARY;
  if (condition == HYB_PRED_IS_NULL || condition == HYB_PRED_IS_NOT_NULL) {
    scan_is_null(segment, chunk_id, condition, position_filter, matches);
  } else if (is_between_condition(condition)) {
    if (dictionary) {
      const uint32_t* bounds = predicate->value_id_bounds ? predicate->value_id_bounds + 4 * size_t{chunk_id} : nullptr;
      scan_dictionary_between(segment, chunk_id, condition, predicate->lower, predicate->upper, bounds, position_filter,
                              matches);
    } else {
      scan_generic_between(segment, chunk_id, condition, predicate->lower, predicate->upper, position_filter, matches);
    }
  } else {
    if (dictionary) {
      const uint32_t* bounds = predicate->value_id_bounds ? predicate->value_id_bounds + 2 * size_t{chunk_id} : nullptr;
      scan_dictionary_vs_value(segment, chunk_id, condition, predicate->lower, bounds, position_filter, matches);
    } else {
      scan_generic_vs_value(segment, chunk_id, condition, predicate->lower, position_filter, matches);
    }
  }
}

int check_predicate(const hyb_table_view* table, const hyb_scan_predicate* predicate) {
  ORC_CHECK(predicate->column_id < table->column_count, HYB_ERR_INVALID, "predicate column out of range");
  const int32_t c = predicate->condition;
  ORC_CHECK((c >= HYB_PRED_EQUALS && c <= HYB_PRED_BETWEEN_EXCLUSIVE) || c == HYB_PRED_IS_NULL ||
                c == HYB_PRED_IS_NOT_NULL,
            HYB_ERR_UNSUPPORTED, "predicate condition not restated");
  return HYB_OK;
}

// TableScan::_on_execute (table_scan.cpp:97-240): one job per chunk; for reference-table input the matches are mapped
// back to the referenced RowIDs (:186-190).
int table_scan(const hyb_table_view* table, const hyb_scan_predicate* predicate, const orc_pos_list* input_filter,
               int threads, std::vector<RowIDs>& per_chunk) {
  if (int status = check_predicate(table, predicate)) return status;
  if (input_filter) {
    ORC_CHECK(input_filter->chunk_count == table->chunk_count, HYB_ERR_INVALID, "filter/table chunk count mismatch");
  }
  per_chunk.assign(table->chunk_count, RowIDs{});
  run_jobs(table->chunk_count, threads, [&](size_t chunk) {
    const uint32_t chunk_id = static_cast<uint32_t>(chunk);
    if (!input_filter) {
      scan_chunk(table, chunk_id, predicate, nullptr, per_chunk[chunk]);
      return;
    }
    const uint64_t begin = input_filter->chunk_offsets[chunk], end = input_filter->chunk_offsets[chunk + 1];
    if (begin == end) return;  // such a chunk does not exist in the reference's input table
    const RowIDs pos_list_in(input_filter->row_ids + begin, input_filter->row_ids + end);
    RowIDs matches;
    scan_chunk(table, chunk_id, predicate, &pos_list_in, matches);
    auto& out = per_chunk[chunk];
    out.reserve(matches.size());
    for (const auto& match : matches) out.push_back(pos_list_in[match.chunk_offset]);
  });
  return HYB_OK;
}

void to_pos_list(const std::vector<RowIDs>& per_chunk, orc_pos_list* out) {
  out->chunk_count = static_cast<uint32_t>(per_chunk.size());
  out->chunk_offsets = static_cast<uint64_t*>(std::malloc(sizeof(uint64_t) * (per_chunk.size() + 1)));
  uint64_t total = 0;
  for (size_t chunk = 0; chunk < per_chunk.size(); ++chunk) {
    out->chunk_offsets[chunk] = total;
    total += per_chunk[chunk].size();
  }
  out->chunk_offsets[per_chunk.size()] = total;
  out->total = total;
  out->row_ids = static_cast<hyb_row_id*>(std::malloc(sizeof(hyb_row_id) * std::max<uint64_t>(total, 1)));
  for (size_t chunk = 0; chunk < per_chunk.size(); ++chunk) {
    if (!per_chunk[chunk].empty()) {
      std::memcpy(out->row_ids + out->chunk_offsets[chunk], per_chunk[chunk].data(),
                  sizeof(hyb_row_id) * per_chunk[chunk].size());
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// JoinHash
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint32_t BLOOM_FILTER_SIZE = 1u << 20;  // join_hash_steps.hpp:252
constexpr uint32_t BLOOM_FILTER_MASK = BLOOM_FILTER_SIZE - 1;
constexpr size_t PROBE_SIZE_PER_CHUNK = size_t{HYB_DEFAULT_CHUNK_SIZE} * 2;  // :47
constexpr uint32_t JOB_SPAWN_THRESHOLD = 500;

using BloomFilter = std::vector<uint8_t>;  // one byte per slot (boost::dynamic_bitset in the reference)

template <typename T>
struct PartitionedElement {  // join_hash_steps.hpp:53-57
  hyb_row_id row_id;
  T value;
};

template <typename T>
struct Partition {  // :66-78
  std::vector<PartitionedElement<T>> elements;
  std::vector<uint8_t> null_values;
};

template <typename T>
using RadixContainer = std::vector<Partition<T>>;

inline bool row_id_is_null(const hyb_row_id& row_id) { return row_id.chunk_offset == HYB_INVALID_CHUNK_OFFSET; }
constexpr hyb_row_id NULL_ROW_ID{HYB_INVALID_CHUNK_ID, HYB_INVALID_CHUNK_OFFSET};

// std::hash<int32_t/int64_t> under libstdc++ is the identity cast to size_t (pinned by
// src/test/lib/operators/join_hash/join_hash_steps_test.cpp:169-188).
template <typename HashedType>
inline size_t std_hash(HashedType value) {
  return static_cast<size_t>(value);
}

// materialize_input (join_hash_steps.hpp:274-420)
template <typename T, typename HashedType, bool keep_null_values>
RadixContainer<T> materialize_input(const hyb_table_view* table, uint32_t column_id, const orc_pos_list* filter,
                                    std::vector<std::vector<size_t>>& histograms, size_t radix_bits,
                                    BloomFilter& output_bloom_filter, const BloomFilter* input_bloom_filter, int threads) {
  const uint32_t chunk_count = table->chunk_count;
  RadixContainer<T> radix_container(chunk_count);
  const size_t num_radix_partitions = size_t{1} << radix_bits;
  const size_t radix_mask = num_radix_partitions - 1;
  output_bloom_filter.assign(BLOOM_FILTER_SIZE, 0);
  std::mutex bloom_mutex;
  histograms.assign(chunk_count, {});

  run_jobs(chunk_count, threads, [&](size_t chunk) {
    const uint32_t chunk_id = static_cast<uint32_t>(chunk);
    const auto& segment = segment_at(table, chunk_id, column_id);
    BloomFilter local_bloom(BLOOM_FILTER_SIZE, 0);
    auto& elements = radix_container[chunk].elements;
    auto& null_values = radix_container[chunk].null_values;
    std::vector<size_t> histogram(num_radix_partitions);
    const uint64_t begin = filter ? filter->chunk_offsets[chunk] : 0;
    const uint32_t num_rows = filter ? static_cast<uint32_t>(filter->chunk_offsets[chunk + 1] - begin) : segment.row_count;
    elements.reserve(num_rows);
    for (uint32_t position = 0; position < num_rows; ++position) {
      // Reference-table input: the value comes from the referenced row; the emitted RowID is the one that
      // write_output_chunks finally dereferences to (join_output_writing.cpp:127-149).
      const hyb_row_id row_id = filter ? filter->row_ids[begin + position] : hyb_row_id{chunk_id, position};
      const uint32_t offset = row_id.chunk_offset;
      const bool is_null = segment_is_null(segment, offset);
      if (!is_null || keep_null_values) {
        const T value = segment_value<T>(segment, offset);
        const size_t hashed_value = std_hash<HashedType>(static_cast<HashedType>(value));
        bool skip = false;
        if (!is_null && input_bloom_filter && !(*input_bloom_filter)[hashed_value & BLOOM_FILTER_MASK] &&
            !keep_null_values) {
          skip = true;
        }
        if (!skip) {
          local_bloom[hashed_value & BLOOM_FILTER_MASK] = 1;
          elements.push_back(PartitionedElement<T>{row_id, value});
          if constexpr (keep_null_values) null_values.push_back(is_null ? 1 : 0);
          if (radix_bits > 0) ++histogram[hashed_value & radix_mask];
        }
      }
    }
    histograms[chunk] = std::move(histogram);
    std::lock_guard<std::mutex> lock(bloom_mutex);
    for (uint32_t slot = 0; slot < BLOOM_FILTER_SIZE; ++slot) output_bloom_filter[slot] |= local_bloom[slot];
  });
  if (filter) {
    // A reference-table input only has the chunks its producer emitted: TableScan drops chunks without matches
    // (table_scan.cpp:132-134), so they do not exist as partitions here either (matters for build(): an input with
    // zero chunks yields no hash table at all, join_hash_steps.hpp:433-435).
    RadixContainer<T> compacted;
    std::vector<std::vector<size_t>> compacted_histograms;
    for (uint32_t chunk = 0; chunk < chunk_count; ++chunk) {
      if (filter->chunk_offsets[chunk + 1] == filter->chunk_offsets[chunk]) continue;
      compacted.push_back(std::move(radix_container[chunk]));
      compacted_histograms.push_back(std::move(histograms[chunk]));
    }
    radix_container = std::move(compacted);
    histograms = std::move(compacted_histograms);
  }
  return radix_container;
}

// partition_by_radix (join_hash_steps.hpp:509-617): stable scatter; order inside a partition = chunk order, row order.
template <typename T, typename HashedType, bool keep_null_values>
RadixContainer<T> partition_by_radix(const RadixContainer<T>& radix_container,
                                     std::vector<std::vector<size_t>>& histograms, size_t radix_bits, int threads) {
  if (radix_container.empty()) return radix_container;
  const size_t input_partition_count = radix_container.size();
  const size_t output_partition_count = size_t{1} << radix_bits;
  const size_t radix_mask = output_partition_count - 1;
  RadixContainer<T> output(output_partition_count);
  std::vector<std::vector<size_t>> output_offsets(input_partition_count, std::vector<size_t>(output_partition_count));
  for (size_t out_idx = 0; out_idx < output_partition_count; ++out_idx) {
    size_t size = 0;
    for (size_t in_idx = 0; in_idx < input_partition_count; ++in_idx) {
      output_offsets[in_idx][out_idx] = size;
      size += histograms[in_idx][out_idx];
    }
    output[out_idx].elements.resize(size);
    if (keep_null_values) output[out_idx].null_values.resize(size);
  }
  run_jobs(input_partition_count, threads, [&](size_t in_idx) {
    const auto& input = radix_container[in_idx];
    for (size_t index = 0; index < input.elements.size(); ++index) {
      const auto& element = input.elements[index];
      const size_t radix = std_hash<HashedType>(static_cast<HashedType>(element.value)) & radix_mask;
      size_t& out_idx = output_offsets[in_idx][radix];
      if constexpr (keep_null_values) output[radix].null_values[out_idx] = input.null_values[index];
      output[radix].elements[out_idx] = element;
      ++out_idx;
    }
  });
  return output;
}

// PosHashTable (join_hash_steps.hpp:97-236): key -> dense offset in insertion order, positions per offset in insertion
// order, finalize() -> CSR.
template <typename HashedType>
class PosHashTable {
 public:
  explicit PosHashTable(bool all_positions, size_t max_size) : _all_positions(all_positions) {
    _offsets.reserve(max_size);
    if (all_positions) _small_pos_lists.reserve(max_size);
  }
  template <typename InputType>
  void emplace(const InputType& value, hyb_row_id row_id) {
    const auto casted = static_cast<HashedType>(value);
    const auto it = _offsets.emplace(casted, static_cast<uint32_t>(_offsets.size()));
    if (_all_positions) {
      if (it.second) _small_pos_lists.emplace_back();
      _small_pos_lists[it.first->second].push_back(row_id);
    }
  }
  void finalize() {
    if (!_all_positions) return;
    _csr_offsets.resize(_small_pos_lists.size() + 1);
    size_t total = 0;
    for (size_t index = 0; index < _small_pos_lists.size(); ++index) {
      _csr_offsets[index] = total;
      total += _small_pos_lists[index].size();
    }
    _csr_offsets.back() = total;
    _pos_list.reserve(total);
    for (const auto& list : _small_pos_lists) _pos_list.insert(_pos_list.end(), list.begin(), list.end());
    _small_pos_lists.clear();
    _small_pos_lists.shrink_to_fit();
  }
  template <typename InputType>
  std::pair<const hyb_row_id*, const hyb_row_id*> find(const InputType& value) const {
    const auto it = _offsets.find(static_cast<HashedType>(value));
    if (it == _offsets.end()) return {nullptr, nullptr};
    return {_pos_list.data() + _csr_offsets[it->second], _pos_list.data() + _csr_offsets[it->second + 1]};
  }
  template <typename InputType>
  bool contains(const InputType& value) const {
    return _offsets.find(static_cast<HashedType>(value)) != _offsets.end();
  }

 private:
  bool _all_positions;
  std::unordered_map<HashedType, uint32_t> _offsets;
  std::vector<std::vector<hyb_row_id>> _small_pos_lists;
  std::vector<size_t> _csr_offsets;
  RowIDs _pos_list;
};

// build (join_hash_steps.hpp:426-507)
template <typename T, typename HashedType>
std::vector<std::unique_ptr<PosHashTable<HashedType>>> build(const RadixContainer<T>& radix_container,
                                                             bool all_positions, size_t radix_bits,
                                                             const BloomFilter& input_bloom_filter, int threads) {
  std::vector<std::unique_ptr<PosHashTable<HashedType>>> hash_tables;
  if (radix_container.empty()) return hash_tables;
  if (radix_bits == 0) {
    size_t total = 0;
    for (const auto& partition : radix_container) total += partition.elements.size();
    hash_tables.resize(1);
    hash_tables[0] = std::make_unique<PosHashTable<HashedType>>(all_positions, total);
  } else {
    hash_tables.resize(radix_container.size());
  }
  const auto insert = [&](size_t partition_idx) {
    const auto& elements = radix_container[partition_idx].elements;
    if (elements.empty()) return;
    const size_t table_idx = radix_bits > 0 ? partition_idx : 0;
    if (radix_bits > 0) hash_tables[table_idx] = std::make_unique<PosHashTable<HashedType>>(all_positions, elements.size());
    auto& hash_table = *hash_tables[table_idx];
    for (const auto& element : elements) {
      const size_t hashed = std_hash<HashedType>(static_cast<HashedType>(element.value));
      if (!input_bloom_filter[hashed & BLOOM_FILTER_MASK]) continue;
      hash_table.emplace(element.value, element.row_id);
    }
    if (radix_bits > 0) hash_table.finalize();
  };
  if (radix_bits == 0) {
    for (size_t partition_idx = 0; partition_idx < radix_container.size(); ++partition_idx) insert(partition_idx);
    hash_tables[0]->finalize();
  } else {
    run_jobs(radix_container.size(), threads, insert);
  }
  return hash_tables;
}

struct JoinOutput {
  std::vector<RowIDs> build_side;  // one list per (partition, slice)
  std::vector<RowIDs> probe_side;
  std::vector<uint32_t> partition_of_list;
};

struct Slice {
  size_t partition, begin, end, output_idx;
};

template <typename T>
std::vector<Slice> make_slices(const RadixContainer<T>& probe_container, JoinOutput& out, bool with_build_lists) {
  std::vector<Slice> slices;
  for (size_t partition_idx = 0; partition_idx < probe_container.size(); ++partition_idx) {
    const size_t count = probe_container[partition_idx].elements.size();
    if (count == 0) continue;  // skip empty partitions (:644-647)
    for (size_t begin = 0; begin < count; begin += PROBE_SIZE_PER_CHUNK) {
      slices.push_back({partition_idx, begin, std::min(begin + PROBE_SIZE_PER_CHUNK, count), out.probe_side.size()});
      if (with_build_lists) out.build_side.emplace_back();
      out.probe_side.emplace_back();
      out.partition_of_list.push_back(static_cast<uint32_t>(partition_idx));
    }
  }
  return slices;
}

// probe (join_hash_steps.hpp:624-792) without secondary predicates
template <typename T, typename HashedType, bool keep_null_values>
void probe(const RadixContainer<T>& probe_container,
           const std::vector<std::unique_ptr<PosHashTable<HashedType>>>& hash_tables, JoinOutput& out, int32_t mode,
           int threads) {
  const auto slices = make_slices(probe_container, out, true);
  run_jobs(slices.size(), threads, [&](size_t slice_idx) {
    const auto& slice = slices[slice_idx];
    const auto& partition = probe_container[slice.partition];
    const auto& elements = partition.elements;
    RowIDs build_local, probe_local;
    const size_t table_idx = hash_tables.size() > 1 ? slice.partition : 0;
    if (!hash_tables.empty() && hash_tables.at(table_idx)) {
      const auto& hash_table = *hash_tables[table_idx];
      for (size_t offset = slice.begin; offset < slice.end; ++offset) {
        const auto& element = elements[offset];
        if (mode == HYB_JOIN_INNER && row_id_is_null(element.row_id)) continue;
        const auto range = hash_table.find(static_cast<HashedType>(element.value));
        if (range.first != range.second) {
          if constexpr (keep_null_values) {
            if (partition.null_values[offset]) {
              build_local.push_back(NULL_ROW_ID);
              probe_local.push_back(element.row_id);
              continue;
            }
          }
          for (const hyb_row_id* it = range.first; it != range.second; ++it) {
            build_local.push_back(*it);
            probe_local.push_back(element.row_id);
          }
        } else {
          if constexpr (keep_null_values) {
            build_local.push_back(NULL_ROW_ID);
            probe_local.push_back(element.row_id);
          }
        }
      }
    } else {
      if constexpr (keep_null_values) {
        for (size_t offset = slice.begin; offset < slice.end; ++offset) {
          build_local.push_back(NULL_ROW_ID);
          probe_local.push_back(elements[offset].row_id);
        }
      }
    }
    out.build_side[slice.output_idx] = std::move(build_local);
    out.probe_side[slice.output_idx] = std::move(probe_local);
  });
}

// probe_semi_anti (join_hash_steps.hpp:794-922) without secondary predicates
template <typename T, typename HashedType>
void probe_semi_anti(const RadixContainer<T>& probe_container,
                     const std::vector<std::unique_ptr<PosHashTable<HashedType>>>& hash_tables, JoinOutput& out,
                     int32_t mode, bool build_table_is_empty, int threads) {
  const auto slices = make_slices(probe_container, out, false);
  run_jobs(slices.size(), threads, [&](size_t slice_idx) {
    const auto& slice = slices[slice_idx];
    const auto& partition = probe_container[slice.partition];
    const auto& elements = partition.elements;
    const auto& null_values = partition.null_values;
    RowIDs local;
    const size_t table_idx = hash_tables.size() > 1 ? slice.partition : 0;
    if (!hash_tables.empty() && hash_tables.at(table_idx)) {
      const auto& hash_table = *hash_tables[table_idx];
      for (size_t offset = slice.begin; offset < slice.end; ++offset) {
        const auto& element = elements[offset];
        if (mode == HYB_JOIN_SEMI) {
          if (row_id_is_null(element.row_id)) continue;
        } else if (mode == HYB_JOIN_ANTI_NULL_AS_FALSE) {
          if (null_values[offset]) {
            local.push_back(element.row_id);
            continue;
          }
        } else {
          if (null_values[offset]) continue;
        }
        const bool any_match = hash_table.contains(static_cast<HashedType>(element.value));
        if ((mode == HYB_JOIN_SEMI && any_match) || (mode != HYB_JOIN_SEMI && !any_match)) {
          local.push_back(element.row_id);
        }
      }
    } else if (mode == HYB_JOIN_ANTI_NULL_AS_FALSE) {
      for (size_t offset = slice.begin; offset < slice.end; ++offset) local.push_back(elements[offset].row_id);
    } else if (mode == HYB_JOIN_ANTI_NULL_AS_TRUE) {
      for (size_t offset = slice.begin; offset < slice.end; ++offset) {
        if (null_values[offset] && !build_table_is_empty) continue;
        local.push_back(elements[offset].row_id);
      }
    }
    out.probe_side[slice.output_idx] = std::move(local);
  });
}

uint64_t table_row_count(const hyb_table_view* table, const orc_pos_list* filter) {
  if (filter) return filter->total;
  uint64_t rows = 0;
  for (uint32_t chunk = 0; chunk < table->chunk_count; ++chunk) rows += segment_at(table, chunk, 0).row_count;
  return rows;
}

// JoinHashImpl::_on_execute (join_hash.cpp:270-572)
template <typename BuildT, typename ProbeT>
int join_hash_impl(const hyb_table_view* build_table, uint32_t build_column, const orc_pos_list* build_filter,
                   const hyb_table_view* probe_table, uint32_t probe_column, const orc_pos_list* probe_filter,
                   int32_t mode, size_t radix_bits, int threads, orc_join_result* result) {
  using HashedType = std::conditional_t<(sizeof(BuildT) < sizeof(ProbeT)), ProbeT, BuildT>;  // join_hash_traits.hpp:25-31
  const bool keep_nulls_build = mode == HYB_JOIN_ANTI_NULL_AS_TRUE;
  const bool keep_nulls_probe = mode == HYB_JOIN_LEFT || mode == HYB_JOIN_RIGHT ||
                                mode == HYB_JOIN_ANTI_NULL_AS_TRUE || mode == HYB_JOIN_ANTI_NULL_AS_FALSE;
  std::vector<std::vector<size_t>> histograms_build, histograms_probe;
  RadixContainer<BuildT> materialized_build;
  RadixContainer<ProbeT> materialized_probe;
  BloomFilter build_bloom, probe_bloom;

  const auto materialize_build = [&](const BloomFilter* input) {
    materialized_build =
        keep_nulls_build
            ? materialize_input<BuildT, HashedType, true>(build_table, build_column, build_filter, histograms_build,
                                                          radix_bits, build_bloom, input, threads)
            : materialize_input<BuildT, HashedType, false>(build_table, build_column, build_filter, histograms_build,
                                                           radix_bits, build_bloom, input, threads);
  };
  const auto materialize_probe = [&](const BloomFilter* input) {
    materialized_probe =
        keep_nulls_probe
            ? materialize_input<ProbeT, HashedType, true>(probe_table, probe_column, probe_filter, histograms_probe,
                                                          radix_bits, probe_bloom, input, threads)
            : materialize_input<ProbeT, HashedType, false>(probe_table, probe_column, probe_filter, histograms_probe,
                                                           radix_bits, probe_bloom, input, threads);
  };
  // (:364-381) the smaller side is materialized first; its Bloom filter prunes the other.
  if (table_row_count(build_table, build_filter) < table_row_count(probe_table, probe_filter)) {
    materialize_build(nullptr);
    materialize_probe(&build_bloom);
  } else {
    materialize_probe(nullptr);
    materialize_build(&probe_bloom);
  }
  for (const auto& partition : materialized_build) result->build_materialized += partition.elements.size();
  for (const auto& partition : materialized_probe) result->probe_materialized += partition.elements.size();

  RadixContainer<BuildT> radix_build;
  RadixContainer<ProbeT> radix_probe;
  if (radix_bits > 0) {
    radix_build = keep_nulls_build
                      ? partition_by_radix<BuildT, HashedType, true>(materialized_build, histograms_build, radix_bits, threads)
                      : partition_by_radix<BuildT, HashedType, false>(materialized_build, histograms_build, radix_bits, threads);
    radix_probe = keep_nulls_probe
                      ? partition_by_radix<ProbeT, HashedType, true>(materialized_probe, histograms_probe, radix_bits, threads)
                      : partition_by_radix<ProbeT, HashedType, false>(materialized_probe, histograms_probe, radix_bits, threads);
    materialized_build.clear();
    materialized_probe.clear();
  } else {
    radix_build = std::move(materialized_build);
    radix_probe = std::move(materialized_probe);
  }

  const bool semi_or_anti =
      mode == HYB_JOIN_SEMI || mode == HYB_JOIN_ANTI_NULL_AS_TRUE || mode == HYB_JOIN_ANTI_NULL_AS_FALSE;
  auto hash_tables = build<BuildT, HashedType>(radix_build, /*all_positions=*/!semi_or_anti, radix_bits, probe_bloom, threads);

  JoinOutput output;
  bool early_out = false;
  if (mode == HYB_JOIN_ANTI_NULL_AS_TRUE) {  // (:471-483)
    for (const auto& partition : radix_build) {
      for (const auto null_value : partition.null_values) {
        if (null_value) early_out = true;
      }
    }
  }
  if (!early_out) {
    switch (mode) {
      case HYB_JOIN_INNER:
        probe<ProbeT, HashedType, false>(radix_probe, hash_tables, output, mode, threads);
        break;
      case HYB_JOIN_LEFT:
      case HYB_JOIN_RIGHT:
        probe<ProbeT, HashedType, true>(radix_probe, hash_tables, output, mode, threads);
        break;
      default:
        probe_semi_anti<ProbeT, HashedType>(radix_probe, hash_tables, output, mode,
                                            table_row_count(build_table, build_filter) == 0, threads);
        break;
    }
  }

  // Flatten. partition_offsets: by radix partition; slice_offsets: the pos lists as probe() produced them.
  const size_t list_count = output.probe_side.size();
  uint64_t total = 0;
  for (const auto& list : output.probe_side) total += list.size();
  result->pair_count = total;
  result->radix_bits = static_cast<int32_t>(radix_bits);
  result->partition_count = static_cast<uint32_t>(size_t{1} << radix_bits);
  result->probe_row_ids = static_cast<hyb_row_id*>(std::malloc(sizeof(hyb_row_id) * std::max<uint64_t>(total, 1)));
  result->build_row_ids =
      semi_or_anti ? nullptr : static_cast<hyb_row_id*>(std::malloc(sizeof(hyb_row_id) * std::max<uint64_t>(total, 1)));
  result->slice_count = static_cast<uint32_t>(list_count);
  result->slice_offsets = static_cast<uint64_t*>(std::malloc(sizeof(uint64_t) * (list_count + 1)));
  result->partition_offsets = static_cast<uint64_t*>(std::malloc(sizeof(uint64_t) * (result->partition_count + 1)));
  std::vector<uint64_t> partition_sizes(result->partition_count, 0);
  uint64_t cursor = 0;
  for (size_t list = 0; list < list_count; ++list) {
    result->slice_offsets[list] = cursor;
    const auto& probe_list = output.probe_side[list];
    if (!probe_list.empty()) {
      std::memcpy(result->probe_row_ids + cursor, probe_list.data(), sizeof(hyb_row_id) * probe_list.size());
      if (!semi_or_anti) {
        std::memcpy(result->build_row_ids + cursor, output.build_side[list].data(), sizeof(hyb_row_id) * probe_list.size());
      }
    }
    // With radix_bits == 0 the "partitions" of the probe container are input chunks: all belong to partition 0.
    partition_sizes[radix_bits > 0 ? output.partition_of_list[list] : 0] += probe_list.size();
    cursor += probe_list.size();
  }
  result->slice_offsets[list_count] = cursor;
  uint64_t running = 0;
  for (uint32_t partition = 0; partition < result->partition_count; ++partition) {
    result->partition_offsets[partition] = running;
    running += partition_sizes[partition];
  }
  result->partition_offsets[result->partition_count] = running;

  // write_output_chunks merging (join_output_writing.cpp:247-296): MIN_SIZE 1000, MAX_SIZE 4000.
  std::vector<uint64_t> chunk_offsets;
  {
    constexpr size_t MIN_SIZE = 1000, MAX_SIZE = MIN_SIZE * 4;
    size_t partition_id = 0;
    uint64_t position = 0;
    while (partition_id < list_count) {
      size_t size = output.probe_side[partition_id].size();
      if (size == 0) {
        ++partition_id;
        continue;
      }
      chunk_offsets.push_back(position);
      while (partition_id + 1 < list_count && size < MIN_SIZE &&
             size + output.probe_side[partition_id + 1].size() < MAX_SIZE) {
        size += output.probe_side[partition_id + 1].size();
        ++partition_id;
      }
      position += size;
      ++partition_id;
    }
    chunk_offsets.push_back(position);
  }
  result->output_chunk_count = static_cast<uint32_t>(chunk_offsets.size() - 1);
  result->output_chunk_offsets = static_cast<uint64_t*>(std::malloc(sizeof(uint64_t) * chunk_offsets.size()));
  std::memcpy(result->output_chunk_offsets, chunk_offsets.data(), sizeof(uint64_t) * chunk_offsets.size());
  return HYB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Projection arithmetic (expression/evaluation/expression_functors.hpp:128-212, expression_utils.cpp:172-205)
// ---------------------------------------------------------------------------------------------------------------------
int32_t expression_common_type(int32_t lhs, int32_t rhs) {
  if (lhs == HYB_TYPE_FLOAT64 || rhs == HYB_TYPE_FLOAT64) return HYB_TYPE_FLOAT64;
  const auto is_float = [](int32_t t) { return t == HYB_TYPE_FLOAT32 || t == HYB_TYPE_FLOAT64; };
  if (lhs == HYB_TYPE_INT64) return is_float(rhs) ? HYB_TYPE_FLOAT64 : HYB_TYPE_INT64;
  if (rhs == HYB_TYPE_INT64) return is_float(lhs) ? HYB_TYPE_FLOAT64 : HYB_TYPE_INT64;
  if (lhs == HYB_TYPE_FLOAT32 || rhs == HYB_TYPE_FLOAT32) return HYB_TYPE_FLOAT32;
  return HYB_TYPE_INT32;
}

struct Scalar {
  int32_t type = HYB_TYPE_INT32;
  bool is_null = false;
  union {
    int32_t i32;
    int64_t i64;
    float f32;
    double f64;
  };
};

template <typename T>
T scalar_as(const Scalar& s) {
  switch (s.type) {
    case HYB_TYPE_INT32:
      return static_cast<T>(s.i32);
    case HYB_TYPE_INT64:
      return static_cast<T>(s.i64);
    case HYB_TYPE_FLOAT32:
      return static_cast<T>(s.f32);
    default:
      return static_cast<T>(s.f64);
  }
}

template <typename T>
void scalar_set(Scalar& s, T value) {
  if constexpr (std::is_same_v<T, int32_t>) {
    s.type = HYB_TYPE_INT32;
    s.i32 = value;
  } else if constexpr (std::is_same_v<T, int64_t>) {
    s.type = HYB_TYPE_INT64;
    s.i64 = value;
  } else if constexpr (std::is_same_v<T, float>) {
    s.type = HYB_TYPE_FLOAT32;
    s.f32 = value;
  } else {
    s.type = HYB_TYPE_FLOAT64;
    s.f64 = value;
  }
}

// STLArithmeticFunctorWrapper: compute in std::common_type_t<A, B>, cast to the Result type (:136-145).
Scalar apply_arithmetic(int32_t op, const Scalar& a, const Scalar& b) {
  Scalar result;
  const int32_t result_type = expression_common_type(a.type, b.type);
  result.type = result_type;
  result.is_null = a.is_null || b.is_null;
  result.i64 = 0;
  with_type(a.type, [&](auto tag_a) {
    using A = decltype(tag_a);
    return with_type(b.type, [&](auto tag_b) {
      using B = decltype(tag_b);
      using Common = std::common_type_t<A, B>;
      return with_type(result_type, [&](auto tag_r) {
        using R = decltype(tag_r);
        const A va = scalar_as<A>(a);
        const B vb = scalar_as<B>(b);
        if (op == HYB_EXPR_DIV) {  // DivisionEvaluator (:188-212)
          if (!result.is_null) {
            if (vb == 0) {
              result.is_null = true;
            } else {
              scalar_set<R>(result, static_cast<R>(static_cast<R>(va) / static_cast<R>(vb)));
            }
          }
          return 0;
        }
        if (result.is_null) return 0;
        Common value{};
        const Common ca = static_cast<Common>(va), cb = static_cast<Common>(vb);
        if (op == HYB_EXPR_ADD) value = ca + cb;
        if (op == HYB_EXPR_SUB) value = ca - cb;
        if (op == HYB_EXPR_MUL) value = ca * cb;
        scalar_set<R>(result, static_cast<R>(value));
        return 0;
      });
    });
  });
  result.type = result_type;
  return result;
}

int expression_result_type(const hyb_table_view* table, const hyb_aggregate_def& def, int32_t* out_type) {
  std::vector<int32_t> stack;
  for (uint32_t n = 0; n < def.node_count; ++n) {
    const auto& node = def.nodes[n];
    if (node.op == HYB_EXPR_COLUMN) {
      ORC_CHECK(node.column_id < table->column_count, HYB_ERR_INVALID, "expression column out of range");
      ORC_CHECK(table->chunk_count > 0, HYB_ERR_INVALID, "cannot type an expression on a table without chunks");
      stack.push_back(segment_at(table, 0, node.column_id).data_type);
    } else if (node.op == HYB_EXPR_LITERAL) {
      stack.push_back(node.literal_type);
    } else {
      ORC_CHECK(stack.size() >= 2, HYB_ERR_INVALID, "malformed expression");
      const int32_t b = stack.back();
      stack.pop_back();
      const int32_t a = stack.back();
      stack.pop_back();
      stack.push_back(expression_common_type(a, b));
    }
  }
  ORC_CHECK(stack.size() == 1, HYB_ERR_INVALID, "malformed expression");
  *out_type = stack[0];
  return HYB_OK;
}

Scalar evaluate_expression(const hyb_table_view* table, const hyb_aggregate_def& def, hyb_row_id row) {
  Scalar stack[HYB_MAX_EXPR_NODES];
  int top = 0;
  for (uint32_t n = 0; n < def.node_count; ++n) {
    const auto& node = def.nodes[n];
    if (node.op == HYB_EXPR_COLUMN) {
      const auto& segment = segment_at(table, row.chunk_id, node.column_id);
      Scalar s;
      s.is_null = segment_is_null(segment, row.chunk_offset);
      with_type(segment.data_type, [&](auto tag) {
        using T = decltype(tag);
        scalar_set<T>(s, segment_value<T>(segment, row.chunk_offset));
        return 0;
      });
      stack[top++] = s;
    } else if (node.op == HYB_EXPR_LITERAL) {
      Scalar s;
      with_type(node.literal_type, [&](auto tag) {
        using T = decltype(tag);
        scalar_set<T>(s, value_of<T>(node.literal));
        return 0;
      });
      stack[top++] = s;
    } else {
      const Scalar b = stack[--top];
      const Scalar a = stack[--top];
      stack[top++] = apply_arithmetic(node.op, a, b);
    }
  }
  return stack[0];
}

// ---------------------------------------------------------------------------------------------------------------------
// AggregateHash
// ---------------------------------------------------------------------------------------------------------------------
using AggregateKeyEntry = uint64_t;
constexpr AggregateKeyEntry CACHE_MASK = AggregateKeyEntry{1} << 63;  // aggregate_hash.cpp:315
constexpr size_t MAX_KEY_COLUMNS = HYB_MAX_GROUPBY_COLUMNS;

struct AggregateKey {
  std::array<AggregateKeyEntry, MAX_KEY_COLUMNS> entries{};
  bool operator==(const AggregateKey& other) const { return entries == other.entries; }
};
struct AggregateKeyHash {
  size_t operator()(const AggregateKey& key) const {
    size_t seed = 0;
    for (const auto entry : key.entries) seed ^= std::hash<uint64_t>{}(entry) + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
    return seed;
  }
};

struct AggregateResultEntry {  // aggregate_hash.hpp AggregateResult
  double acc_f = 0.0;          // SUM/AVG over float/double (and AVG over ints): double accumulator
  int64_t acc_i = 0;           // SUM over ints
  Scalar extreme;              // MIN/MAX current value
  size_t aggregate_count = 0;
  hyb_row_id row_id = NULL_ROW_ID;
};

struct AggregateContext {
  std::unordered_map<AggregateKey, size_t, AggregateKeyHash> result_ids;
  std::vector<AggregateResultEntry> results;
};

bool scalar_less(const Scalar& a, const Scalar& b) {
  return with_type(a.type, [&](auto tag) {
    using T = decltype(tag);
    return scalar_as<T>(a) < scalar_as<T>(b);
  });
}

}  // namespace

// =====================================================================================================================
// C interface
// =====================================================================================================================
// Predicate normalisation: lossless_cast (lossless_cast.hpp:31-176), next_float_towards / lossless_predicate_cast
// (utils/lossless_predicate_cast.{hpp:20-62,cpp:14-38}), flip / between helpers (types.cpp:51-153) — restated with the
// reference's template structure: one overload set per (Source, Target) pair, std::optional for "not lossless".
// =====================================================================================================================
namespace predicate_cast {

template <typename Target, typename Source>
std::optional<Target> lossless(const Source& source) {
  if constexpr (std::is_same_v<Target, Source>) {
    return source;                                                              // identity (:31-36)
  } else if constexpr (std::is_same_v<Source, int64_t> && std::is_same_v<Target, int32_t>) {
    if (source < std::numeric_limits<int32_t>::min() || source > std::numeric_limits<int32_t>::max()) return std::nullopt;
    return static_cast<int32_t>(source);                                        // (:38-46)
  } else if constexpr (std::is_same_v<Source, int32_t> && std::is_same_v<Target, int64_t>) {
    return static_cast<int64_t>(source);                                        // (:48-53)
  } else if constexpr (std::is_integral_v<Source> && std::is_floating_point_v<Target>) {
    const auto floating_point = static_cast<Target>(source);                    // (:101-110), with the round trip done in
    if (!(static_cast<long double>(floating_point) == static_cast<long double>(source))) return std::nullopt;  // long double
    return floating_point;                                                      // to stay clear of UB at 2^31 / 2^63
  } else if constexpr (std::is_floating_point_v<Source> && std::is_integral_v<Target>) {
    Source integral_part{};                                                     // (:113-147)
    if (std::modf(source, &integral_part) != Source{}) return std::nullopt;
    if (!std::isfinite(source)) return std::nullopt;
    if constexpr (std::is_same_v<Source, float> && std::is_same_v<Target, int32_t>) {
      if (source >= 2'147'483'648.0f || source <= -2'147'483'904.0f) return std::nullopt;
    } else if constexpr (std::is_same_v<Source, double> && std::is_same_v<Target, int32_t>) {
      if (source >= 2'147'483'648.0 || source <= -2'147'483'649.0) return std::nullopt;
    } else if constexpr (std::is_same_v<Source, float> && std::is_same_v<Target, int64_t>) {
      if (source >= 9'223'372'036'854'775'808.0f || source <= -9'223'373'136'366'403'584.0f) return std::nullopt;
    } else {
      if (source >= 9'223'372'036'854'775'808.0 || source <= -9'223'372'036'854'777'856.0) return std::nullopt;
    }
    return static_cast<Target>(source);
  } else if constexpr (std::is_same_v<Source, float> && std::is_same_v<Target, double>) {
    return static_cast<double>(source);                                         // (:149-154)
  } else {
    static_assert(std::is_same_v<Source, double> && std::is_same_v<Target, float>);
    if (source > 340282346638528859811704183484516925440.0 || source < -340282346638528859811704183484516925440.0) {
      return std::nullopt;                                                      // (:156-170)
    }
    const auto casted_source = static_cast<float>(source);
    if (static_cast<double>(casted_source) == source) return casted_source;
    return std::nullopt;
  }
}

std::optional<float> next_float_towards(const double value, const double towards) {  // lossless_predicate_cast.cpp:14-38
  if (value > 340282346638528859811704183484516925440.0 || value < -340282346638528859811704183484516925440.0) return std::nullopt;
  if (value == towards) return std::nullopt;
  const auto casted_value = static_cast<float>(value);
  if ((static_cast<double>(casted_value) < value && towards < value) || (static_cast<double>(casted_value) > value && towards > value)) {
    return casted_value;
  }
  const float next = std::nexttowardf(casted_value, static_cast<long double>(towards));
  if (!std::isfinite(next)) return std::nullopt;
  return next;
}

template <typename Output, typename Input>
std::optional<std::pair<int32_t, Output>> predicate(const int32_t condition, const Input& input) {  // .hpp:20-62
  if (const auto casted = lossless<Output>(input)) return std::make_pair(condition, *casted);
  if (condition < HYB_PRED_EQUALS || condition > HYB_PRED_GREATER_THAN_EQUALS) return std::nullopt;
  if constexpr (std::is_same_v<Input, double> && std::is_same_v<Output, float>) {
    if (condition == HYB_PRED_EQUALS) return std::nullopt;
    if (condition == HYB_PRED_LESS_THAN || condition == HYB_PRED_LESS_THAN_EQUALS) {
      const auto adjusted = next_float_towards(input, std::numeric_limits<double>::lowest());
      if (!adjusted) return std::nullopt;
      return std::make_pair(int32_t{HYB_PRED_LESS_THAN_EQUALS}, *adjusted);
    }
    if (condition == HYB_PRED_GREATER_THAN || condition == HYB_PRED_GREATER_THAN_EQUALS) {
      const auto adjusted = next_float_towards(input, std::numeric_limits<double>::max());
      if (!adjusted) return std::nullopt;
      return std::make_pair(int32_t{HYB_PRED_GREATER_THAN_EQUALS}, *adjusted);
    }
  }
  return std::nullopt;
}

template <typename Output>
void store(hyb_value* out, Output value) {
  if constexpr (std::is_same_v<Output, int32_t>) out->i32 = value;
  if constexpr (std::is_same_v<Output, int64_t>) out->i64 = value;
  if constexpr (std::is_same_v<Output, float>) out->f32 = value;
  if constexpr (std::is_same_v<Output, double>) out->f64 = value;
}

// lossless_predicate_variant_cast (.cpp:40-73): resolve the variant's type and the target type
template <typename Input>
bool variant_to(const int32_t condition, const Input& input, const int32_t target_type, int32_t* out_condition, hyb_value* out_value) {
  const auto finish = [&](const auto& result) {
    if (!result) return false;
    *out_condition = result->first;
    store(out_value, result->second);
    return true;
  };
  switch (target_type) {
    case HYB_TYPE_INT32:
      return finish(predicate<int32_t>(condition, input));
    case HYB_TYPE_INT64:
      return finish(predicate<int64_t>(condition, input));
    case HYB_TYPE_FLOAT32:
      return finish(predicate<float>(condition, input));
    case HYB_TYPE_FLOAT64:
      return finish(predicate<double>(condition, input));
    default:
      return false;
  }
}

bool variant(const int32_t condition, const int32_t source_type, const hyb_value& value, const int32_t target_type,
             int32_t* out_condition, hyb_value* out_value) {
  switch (source_type) {
    case HYB_TYPE_INT32:
      return variant_to(condition, value.i32, target_type, out_condition, out_value);
    case HYB_TYPE_INT64:
      return variant_to(condition, value.i64, target_type, out_condition, out_value);
    case HYB_TYPE_FLOAT32:
      return variant_to(condition, value.f32, target_type, out_condition, out_value);
    case HYB_TYPE_FLOAT64:
      return variant_to(condition, value.f64, target_type, out_condition, out_value);
    default:
      return false;
  }
}

int32_t flip(const int32_t condition) {  // types.cpp:51-82 (-1: Fail("Can't flip ..."))
  switch (condition) {
    case HYB_PRED_EQUALS:
    case HYB_PRED_NOT_EQUALS:
      return condition;
    case HYB_PRED_LESS_THAN:
      return HYB_PRED_GREATER_THAN;
    case HYB_PRED_LESS_THAN_EQUALS:
      return HYB_PRED_GREATER_THAN_EQUALS;
    case HYB_PRED_GREATER_THAN:
      return HYB_PRED_LESS_THAN;
    case HYB_PRED_GREATER_THAN_EQUALS:
      return HYB_PRED_LESS_THAN_EQUALS;
    default:
      return -1;
  }
}

}  // namespace predicate_cast

// =====================================================================================================================
extern "C" {

const char* orc_last_error(void) { return g_error.c_str(); }

int orc_next_float_towards(double value, double towards, float* out_value) {
  const auto result = predicate_cast::next_float_towards(value, towards);
  if (result) *out_value = *result;
  return result ? 1 : 0;
}

// `column <condition> literal` (value_on_left == 0) or `literal <condition> column` (table_scan.cpp:340-366, :388-390).
// Returns 1 and the condition / value ColumnVsValueTableScanImpl receives, 0 for "ExpressionEvaluator fallback".
int orc_normalize_predicate(int32_t condition, int32_t literal_type, hyb_value literal, int32_t column_type, int32_t value_on_left,
                            int32_t* out_condition, hyb_value* out_value) {
  if (value_on_left) {
    const int32_t flipped = predicate_cast::flip(condition);
    if (flipped < 0) return 0;
    int32_t adjusted = 0;
    if (!predicate_cast::variant(flipped, literal_type, literal, column_type, &adjusted, out_value)) return 0;
    // predicate_condition = flip(adjusted) (:350), then ColumnVsValueTableScanImpl(..., flip(predicate_condition), ...) (:389)
    *out_condition = predicate_cast::flip(predicate_cast::flip(adjusted));
    return 1;
  }
  return predicate_cast::variant(condition, literal_type, literal, column_type, out_condition, out_value) ? 1 : 0;
}

// BetweenExpression (table_scan.cpp:399-441): split, cast both bounds, reassemble.
int orc_normalize_between(int32_t condition, int32_t lower_type, hyb_value lower, int32_t upper_type, hyb_value upper,
                          int32_t column_type, int32_t* out_condition, hyb_value* out_lower, hyb_value* out_upper) {
  int32_t lower_condition, upper_condition;
  switch (condition) {  // between_to_conditions (types.cpp:119-132)
    case HYB_PRED_BETWEEN_INCLUSIVE:
      lower_condition = HYB_PRED_GREATER_THAN_EQUALS, upper_condition = HYB_PRED_LESS_THAN_EQUALS;
      break;
    case HYB_PRED_BETWEEN_LOWER_EXCLUSIVE:
      lower_condition = HYB_PRED_GREATER_THAN, upper_condition = HYB_PRED_LESS_THAN_EQUALS;
      break;
    case HYB_PRED_BETWEEN_UPPER_EXCLUSIVE:
      lower_condition = HYB_PRED_GREATER_THAN_EQUALS, upper_condition = HYB_PRED_LESS_THAN;
      break;
    case HYB_PRED_BETWEEN_EXCLUSIVE:
      lower_condition = HYB_PRED_GREATER_THAN, upper_condition = HYB_PRED_LESS_THAN;
      break;
    default:
      return 0;
  }
  if (!predicate_cast::variant(lower_condition, lower_type, lower, column_type, &lower_condition, out_lower)) return 0;
  if (!predicate_cast::variant(upper_condition, upper_type, upper, column_type, &upper_condition, out_upper)) return 0;
  // conditions_to_between (types.cpp:134-153)
  if (lower_condition == HYB_PRED_GREATER_THAN) {
    if (upper_condition == HYB_PRED_LESS_THAN) *out_condition = HYB_PRED_BETWEEN_EXCLUSIVE;
    else if (upper_condition == HYB_PRED_LESS_THAN_EQUALS) *out_condition = HYB_PRED_BETWEEN_LOWER_EXCLUSIVE;
    else return 0;
  } else if (lower_condition == HYB_PRED_GREATER_THAN_EQUALS) {
    if (upper_condition == HYB_PRED_LESS_THAN) *out_condition = HYB_PRED_BETWEEN_UPPER_EXCLUSIVE;
    else if (upper_condition == HYB_PRED_LESS_THAN_EQUALS) *out_condition = HYB_PRED_BETWEEN_INCLUSIVE;
    else return 0;
  } else {
    return 0;
  }
  return 1;
}

// DictionaryEncoder::on_encode (dictionary_encoder.hpp:33-103)
int orc_encode_dictionary(int32_t data_type, const void* values, const uint8_t* nulls, uint32_t n, void* out_dictionary,
                          uint32_t* out_dictionary_size, uint32_t* out_value_ids) {
  ORC_CHECK(data_type >= HYB_TYPE_INT32 && data_type <= HYB_TYPE_FLOAT64, HYB_ERR_UNSUPPORTED,
            "numeric dictionaries only");
  with_type(data_type, [&](auto tag) {
    using T = decltype(tag);
    const T* input = static_cast<const T*>(values);
    std::vector<T> dense_values;
    dense_values.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
      if (!(nulls && nulls[i])) dense_values.push_back(input[i]);
    }
    std::vector<T> dictionary(dense_values);
    std::sort(dictionary.begin(), dictionary.end());
    dictionary.erase(std::unique(dictionary.begin(), dictionary.end()), dictionary.end());
    const uint32_t null_value_id = static_cast<uint32_t>(dictionary.size());
    auto values_iter = dense_values.cbegin();
    for (uint32_t i = 0; i < n; ++i) {
      if (!(nulls && nulls[i])) {
        out_value_ids[i] = static_cast<uint32_t>(
            std::distance(dictionary.cbegin(), std::lower_bound(dictionary.cbegin(), dictionary.cend(), *values_iter)));
        ++values_iter;
      } else {
        out_value_ids[i] = null_value_id;
      }
    }
    std::memcpy(out_dictionary, dictionary.data(), sizeof(T) * dictionary.size());
    *out_dictionary_size = null_value_id;
    return 0;
  });
  return HYB_OK;
}

// FixedWidthIntegerCompressor::_compress_using_max_value (fixed_width_integer_compressor.cpp:33-44)
int orc_compress_fixed_width(const uint32_t* in, uint32_t n, uint32_t max_value, void* out, int32_t* out_vector_type) {
  if (max_value <= std::numeric_limits<uint8_t>::max()) {
    for (uint32_t i = 0; i < n; ++i) static_cast<uint8_t*>(out)[i] = static_cast<uint8_t>(in[i]);
    *out_vector_type = HYB_VEC_FIXED_1B;
  } else if (max_value <= std::numeric_limits<uint16_t>::max()) {
    for (uint32_t i = 0; i < n; ++i) static_cast<uint16_t*>(out)[i] = static_cast<uint16_t>(in[i]);
    *out_vector_type = HYB_VEC_FIXED_2B;
  } else {
    for (uint32_t i = 0; i < n; ++i) static_cast<uint32_t*>(out)[i] = in[i];
    *out_vector_type = HYB_VEC_FIXED_4B;
  }
  return HYB_OK;
}

// BitPackingCompressor::compress (bitpacking_compressor.cpp:21-53); layout per compact_iterator.hpp:218-252
int orc_compress_bitpacking(const uint32_t* in, uint32_t n, uint64_t* out_words, int32_t* out_bit_width) {
  size_t required_bits = 1;
  if (n > 0) {
    const uint32_t max_element = *std::max_element(in, in + n);
    if (max_element != 0) required_bits = static_cast<size_t>(std::ceil(std::log2(static_cast<double>(max_element) + 1.0)));
  }
  const size_t words = (static_cast<size_t>(n) * required_bits + 63) / 64;
  std::fill_n(out_words, words, uint64_t{0});
  for (uint32_t i = 0; i < n; ++i) {
    const uint64_t bit = static_cast<uint64_t>(i) * required_bits;
    const uint64_t word = bit / 64;
    const uint32_t shift = static_cast<uint32_t>(bit % 64);
    out_words[word] |= static_cast<uint64_t>(in[i]) << shift;
    if (shift + required_bits > 64) out_words[word + 1] |= static_cast<uint64_t>(in[i]) >> (64 - shift);
  }
  *out_bit_width = static_cast<int32_t>(required_bits);
  return HYB_OK;
}

// FrameOfReferenceEncoder::on_encode (frame_of_reference_encoder.hpp:25-122)
int orc_encode_frame_of_reference(const int32_t* values, const uint8_t* nulls, uint32_t n, int32_t* out_minima,
                                  uint32_t* out_offsets, uint32_t* out_max_offset, int32_t* out_contains_nulls) {
  uint32_t max_offset = 0;
  bool contains_nulls = false;
  const uint32_t block_size = HYB_FOR_BLOCK_SIZE;
  for (uint32_t block_begin = 0, block = 0; block_begin < n; block_begin += block_size, ++block) {
    const uint32_t block_end = std::min(n, block_begin + block_size);
    int32_t min_value = std::numeric_limits<int32_t>::max();
    int32_t max_value = std::numeric_limits<int32_t>::lowest();
    for (uint32_t i = block_begin; i < block_end; ++i) {
      const bool is_null = nulls && nulls[i];
      contains_nulls |= is_null;
      if (!is_null) {
        min_value = std::min(min_value, values[i]);
        max_value = std::max(max_value, values[i]);
      }
    }
    out_minima[block] = min_value;
    for (uint32_t i = block_begin; i < block_end; ++i) {
      int32_t value = (nulls && nulls[i]) ? min_value : values[i];
      const uint32_t offset = static_cast<uint32_t>(value) - static_cast<uint32_t>(min_value);
      out_offsets[i] = offset;
      max_offset = std::max(max_offset, offset);
    }
  }
  *out_max_offset = max_offset;
  *out_contains_nulls = contains_nulls ? 1 : 0;
  return HYB_OK;
}

int orc_decode_segment(const hyb_segment_desc* segment, void* out_values, uint8_t* out_nulls) {
  ORC_CHECK(segment && out_values, HYB_ERR_INVALID, "NULL argument");
  for (uint32_t i = 0; i < segment->row_count; ++i) {
    if (out_nulls) out_nulls[i] = segment_is_null(*segment, i) ? 1 : 0;
  }
  if (segment->data_type == HYB_TYPE_STRING) {
    ORC_CHECK(segment->encoding == HYB_ENC_DICTIONARY, HYB_ERR_UNSUPPORTED, "string segments must be dictionary encoded");
    for (uint32_t i = 0; i < segment->row_count; ++i) {
      static_cast<uint32_t*>(out_values)[i] =
          vector_get(segment->attribute_vector, segment->vector_type, segment->bit_width, i);
    }
    return HYB_OK;
  }
  with_type(segment->data_type, [&](auto tag) {
    using T = decltype(tag);
    for (uint32_t i = 0; i < segment->row_count; ++i) static_cast<T*>(out_values)[i] = segment_value<T>(*segment, i);
    return 0;
  });
  return HYB_OK;
}

int orc_table_scan(const hyb_table_view* table, const hyb_scan_predicate* predicate, const orc_pos_list* input_filter,
                   int32_t threads, orc_pos_list* out) {
  ORC_CHECK(table && predicate && out, HYB_ERR_INVALID, "NULL argument");
  std::vector<RowIDs> per_chunk;
  if (int status = table_scan(table, predicate, input_filter, threads, per_chunk)) return status;
  to_pos_list(per_chunk, out);
  return HYB_OK;
}

void orc_pos_list_free(orc_pos_list* list) {
  if (!list) return;
  std::free(list->chunk_offsets);
  std::free(list->row_ids);
  list->chunk_offsets = nullptr;
  list->row_ids = nullptr;
}

// JoinHash::calculate_radix_bits (join_hash.cpp:70-114)
int32_t orc_calculate_radix_bits(uint64_t build_side_size, uint64_t /*probe_side_size*/) {
  constexpr auto L2_CACHE_SIZE = 1'024'000;
  constexpr auto L2_CACHE_MAX_USABLE = L2_CACHE_SIZE * 0.75;
  const auto complete_hash_map_size = static_cast<double>(build_side_size) * static_cast<double>(sizeof(uint32_t)) / 0.8;
  const auto cluster_count = std::max(1.0, complete_hash_map_size / L2_CACHE_MAX_USABLE);
  return static_cast<int32_t>(std::min(size_t{8}, static_cast<size_t>(std::ceil(std::log2(cluster_count)))));
}

int orc_join_hash(const hyb_table_view* build_table, uint32_t build_column, const orc_pos_list* build_filter,
                  const hyb_table_view* probe_table, uint32_t probe_column, const orc_pos_list* probe_filter,
                  int32_t mode, int32_t radix_bits, int32_t threads, orc_join_result* out) {
  ORC_CHECK(build_table && probe_table && out, HYB_ERR_INVALID, "NULL argument");
  std::memset(out, 0, sizeof(*out));
  ORC_CHECK(build_column < build_table->column_count && probe_column < probe_table->column_count, HYB_ERR_INVALID,
            "join column out of range");
  ORC_CHECK(mode == HYB_JOIN_INNER || mode == HYB_JOIN_LEFT || mode == HYB_JOIN_RIGHT || mode == HYB_JOIN_SEMI ||
                mode == HYB_JOIN_ANTI_NULL_AS_TRUE || mode == HYB_JOIN_ANTI_NULL_AS_FALSE,
            HYB_ERR_UNSUPPORTED, "JoinMode not supported by JoinHash");
  const auto column_type = [](const hyb_table_view* table, uint32_t column) {
    return table->chunk_count ? segment_at(table, 0, column).data_type : HYB_TYPE_INT32;
  };
  const int32_t build_type = column_type(build_table, build_column);
  const int32_t probe_type = column_type(probe_table, probe_column);
  const auto integral = [](int32_t t) { return t == HYB_TYPE_INT32 || t == HYB_TYPE_INT64; };
  ORC_CHECK(integral(build_type) && integral(probe_type), HYB_ERR_UNSUPPORTED, "only integer join keys are restated");
  if (radix_bits < 0) {
    radix_bits = orc_calculate_radix_bits(table_row_count(build_table, build_filter),
                                          table_row_count(probe_table, probe_filter));
  }
  const size_t bits = static_cast<size_t>(radix_bits);
  if (build_type == HYB_TYPE_INT32 && probe_type == HYB_TYPE_INT32) {
    return join_hash_impl<int32_t, int32_t>(build_table, build_column, build_filter, probe_table, probe_column,
                                            probe_filter, mode, bits, threads, out);
  }
  if (build_type == HYB_TYPE_INT32) {
    return join_hash_impl<int32_t, int64_t>(build_table, build_column, build_filter, probe_table, probe_column,
                                            probe_filter, mode, bits, threads, out);
  }
  if (probe_type == HYB_TYPE_INT32) {
    return join_hash_impl<int64_t, int32_t>(build_table, build_column, build_filter, probe_table, probe_column,
                                            probe_filter, mode, bits, threads, out);
  }
  return join_hash_impl<int64_t, int64_t>(build_table, build_column, build_filter, probe_table, probe_column,
                                          probe_filter, mode, bits, threads, out);
}

// Test hook for the KATs of src/test/lib/operators/join_hash/join_hash_steps_test.cpp:169-263: runs materialize_input
// <int32, int32> and reports elements (in container order), per-chunk histograms and the output Bloom filter slots.
int orc_debug_materialize(const hyb_table_view* table, uint32_t column, int32_t keep_nulls, int32_t radix_bits,
                          const uint32_t* input_bloom_slots, uint32_t input_bloom_slot_count, int32_t* out_values,
                          hyb_row_id* out_row_ids, uint8_t* out_nulls, uint64_t* out_count, uint64_t* out_histograms,
                          uint32_t* out_bloom_slots, uint32_t* out_bloom_slot_count) {
  std::vector<std::vector<size_t>> histograms;
  BloomFilter output_bloom, input_bloom;
  const BloomFilter* input = nullptr;
  if (input_bloom_slots) {
    input_bloom.assign(BLOOM_FILTER_SIZE, 0);
    for (uint32_t i = 0; i < input_bloom_slot_count; ++i) input_bloom[input_bloom_slots[i]] = 1;
    input = &input_bloom;
  }
  RadixContainer<int32_t> container =
      keep_nulls ? materialize_input<int32_t, int32_t, true>(table, column, nullptr, histograms, radix_bits, output_bloom, input, 1)
                 : materialize_input<int32_t, int32_t, false>(table, column, nullptr, histograms, radix_bits, output_bloom, input, 1);
  uint64_t count = 0;
  for (const auto& partition : container) {
    for (size_t i = 0; i < partition.elements.size(); ++i) {
      out_values[count] = partition.elements[i].value;
      out_row_ids[count] = partition.elements[i].row_id;
      out_nulls[count] = keep_nulls ? partition.null_values[i] : 0;
      ++count;
    }
  }
  *out_count = count;
  const size_t partitions = size_t{1} << radix_bits;
  for (size_t chunk = 0; chunk < histograms.size(); ++chunk) {
    for (size_t p = 0; p < partitions; ++p) out_histograms[chunk * partitions + p] = histograms[chunk][p];
  }
  uint32_t slots = 0;
  for (uint32_t slot = 0; slot < BLOOM_FILTER_SIZE; ++slot) {
    if (output_bloom[slot] && slots < 4096) out_bloom_slots[slots++] = slot;
  }
  *out_bloom_slot_count = slots;
  return HYB_OK;
}

void orc_join_result_free(orc_join_result* result) {
  if (!result) return;
  std::free(result->build_row_ids);
  std::free(result->probe_row_ids);
  std::free(result->partition_offsets);
  std::free(result->slice_offsets);
  std::free(result->output_chunk_offsets);
  std::memset(result, 0, sizeof(*result));
}

// [TableScan...] -> [Projection] -> AggregateHash::_on_execute (aggregate_hash.cpp:1180-1372)
int orc_aggregate_hash(const hyb_table_view* table, const hyb_aggregate_query* query, const orc_pos_list* filter,
                       int32_t threads, int32_t parallel, orc_aggregate_result* out) {
  ORC_CHECK(table && query && out, HYB_ERR_INVALID, "NULL argument");
  std::memset(out, 0, sizeof(*out));
  ORC_CHECK(query->groupby_count <= MAX_KEY_COLUMNS, HYB_ERR_UNSUPPORTED, "too many group-by columns");
  ORC_CHECK(query->aggregate_count <= HYB_MAX_AGGREGATES, HYB_ERR_UNSUPPORTED, "too many aggregates");
  (void)parallel;

  // 1. The chain of TableScans that precedes the aggregate in the plan: each scan consumes the previous one's output.
  std::vector<RowIDs> rows_per_chunk(table->chunk_count);
  {
    orc_pos_list current{};
    bool have_current = false;
    if (filter) {
      current = *filter;  // borrowed
    }
    const orc_pos_list* input = filter;
    for (uint32_t p = 0; p < query->predicate_count; ++p) {
      orc_pos_list next{};
      if (int status = orc_table_scan(table, &query->predicates[p], input, threads, &next)) {
        if (have_current) orc_pos_list_free(&current);
        return status;
      }
      if (have_current) orc_pos_list_free(&current);
      current = next;
      have_current = true;
      input = &current;
    }
    for (uint32_t chunk = 0; chunk < table->chunk_count; ++chunk) {
      auto& rows = rows_per_chunk[chunk];
      if (input) {
        rows.assign(input->row_ids + input->chunk_offsets[chunk], input->row_ids + input->chunk_offsets[chunk + 1]);
      } else {
        const uint32_t count = segment_at(table, chunk, 0).row_count;
        rows.reserve(count);
        for (uint32_t offset = 0; offset < count; ++offset) rows.push_back(hyb_row_id{chunk, offset});
      }
    }
    if (have_current) orc_pos_list_free(&current);
  }
  // The aggregate's input table holds only non-empty chunks (TableScan drops empty ones, table_scan.cpp:132-134).
  std::vector<const RowIDs*> input_chunks;
  uint64_t input_row_count = 0;
  for (const auto& rows : rows_per_chunk) {
    if (!rows.empty()) {
      input_chunks.push_back(&rows);
      input_row_count += rows.size();
    }
  }

  // 2. Types
  const uint32_t aggregate_count = query->aggregate_count;
  std::vector<int32_t> input_types(aggregate_count, HYB_TYPE_INT64), result_types(aggregate_count, HYB_TYPE_INT64);
  for (uint32_t a = 0; a < aggregate_count; ++a) {
    const auto& def = query->aggregates[a];
    if (def.function == HYB_AGG_COUNT_STAR) {
      result_types[a] = HYB_TYPE_INT64;
      continue;
    }
    ORC_CHECK(def.function >= HYB_AGG_MIN && def.function <= HYB_AGG_COUNT, HYB_ERR_UNSUPPORTED,
              "aggregate function not restated");
    if (int status = expression_result_type(table, def, &input_types[a])) return status;
    ORC_CHECK(input_types[a] != HYB_TYPE_STRING, HYB_ERR_UNSUPPORTED, "aggregates over strings stay on the CPU operator");
    const bool integral = input_types[a] == HYB_TYPE_INT32 || input_types[a] == HYB_TYPE_INT64;
    switch (def.function) {  // window_function_traits.hpp:14-77
      case HYB_AGG_COUNT:
        result_types[a] = HYB_TYPE_INT64;
        break;
      case HYB_AGG_SUM:
        result_types[a] = integral ? HYB_TYPE_INT64 : HYB_TYPE_FLOAT64;
        break;
      case HYB_AGG_AVG:
        result_types[a] = HYB_TYPE_FLOAT64;
        break;
      default:
        result_types[a] = input_types[a];
        break;
    }
  }

  // 3. _partition_by_groupby_keys (aggregate_hash.cpp:661-948): one AggregateKeyEntry per row and group-by column.
  const uint32_t groupby_count = query->groupby_count;
  std::vector<std::vector<AggregateKey>> keys_per_chunk(input_chunks.size());
  for (size_t c = 0; c < input_chunks.size(); ++c) keys_per_chunk[c].resize(input_chunks[c]->size());
  bool use_immediate_key_shortcut = false;
  size_t expected_result_size = 0;
  for (uint32_t g = 0; g < groupby_count; ++g) {
    const uint32_t column_id = query->groupby_column_ids[g];
    ORC_CHECK(column_id < table->column_count, HYB_ERR_INVALID, "group-by column out of range");
    const int32_t data_type = table->chunk_count ? segment_at(table, 0, column_id).data_type : HYB_TYPE_INT32;
    if (data_type == HYB_TYPE_INT32) {
      // (:737-764) value - INT32_MIN + 1; 0 = NULL
      AggregateKeyEntry min_key = std::numeric_limits<AggregateKeyEntry>::max(), max_key = 0;
      for (size_t c = 0; c < input_chunks.size(); ++c) {
        const auto& rows = *input_chunks[c];
        for (size_t i = 0; i < rows.size(); ++i) {
          const auto& segment = segment_at(table, rows[i].chunk_id, column_id);
          if (segment_is_null(segment, rows[i].chunk_offset)) {
            keys_per_chunk[c][i].entries[g] = 0;
          } else {
            const int32_t value = segment_value<int32_t>(segment, rows[i].chunk_offset);
            const auto key = static_cast<uint64_t>(static_cast<int64_t>(value) - std::numeric_limits<int32_t>::min()) + 1;
            keys_per_chunk[c][i].entries[g] = key;
            min_key = std::min(min_key, key);
            max_key = std::max(max_key, key);
          }
        }
      }
      if (groupby_count == 1) {  // immediate key shortcut (:781-804)
        if (max_key > 0 && static_cast<double>(max_key - min_key) < static_cast<double>(input_row_count) * 1.2) {
          expected_result_size = static_cast<size_t>(max_key - min_key) + 2;
          use_immediate_key_shortcut = true;
          for (auto& keys : keys_per_chunk) {
            for (auto& key : keys) {
              auto& entry = key.entries[0];
              entry = entry == 0 ? (entry | CACHE_MASK) : ((entry - min_key + 1) | CACHE_MASK);
            }
          }
        }
      }
    } else if (data_type == HYB_TYPE_STRING) {
      // (:852-925) strings of < 5 chars are packed; longer ones get ids from a map. Only value-IDs reach this code, so
      // the caller supplies the per-dictionary-entry codes it computed with exactly that scheme.
      // one job per chunk, as _partition_by_groupby_keys spawns them (:683-700, :927-940)
      std::atomic<bool> missing_codes{false};
      run_jobs(input_chunks.size(), threads, [&](size_t c) {
        const auto& rows = *input_chunks[c];
        for (size_t i = 0; i < rows.size(); ++i) {
          const auto& segment = segment_at(table, rows[i].chunk_id, column_id);
          if (!(segment.encoding == HYB_ENC_DICTIONARY && segment.dictionary_codes)) {
            missing_codes = true;
            return;
          }
          const uint32_t value_id =
              vector_get(segment.attribute_vector, segment.vector_type, segment.bit_width, rows[i].chunk_offset);
          keys_per_chunk[c][i].entries[g] = value_id >= segment.dictionary_size ? 0 : segment.dictionary_codes[value_id];
        }
      });
      ORC_CHECK(!missing_codes, HYB_ERR_INVALID, "string group-by columns need dictionary_codes");
    } else {
      // (:818-925) dense ids in first-appearance order starting at 1; 0 = NULL
      with_type(data_type, [&](auto tag) {
        using T = decltype(tag);
        std::unordered_map<T, AggregateKeyEntry> id_map;
        AggregateKeyEntry id_counter = 1;
        for (size_t c = 0; c < input_chunks.size(); ++c) {
          const auto& rows = *input_chunks[c];
          for (size_t i = 0; i < rows.size(); ++i) {
            const auto& segment = segment_at(table, rows[i].chunk_id, column_id);
            if (segment_is_null(segment, rows[i].chunk_offset)) {
              keys_per_chunk[c][i].entries[g] = 0;
            } else {
              const auto inserted = id_map.try_emplace(segment_value<T>(segment, rows[i].chunk_offset), id_counter);
              if (inserted.second) ++id_counter;
              keys_per_chunk[c][i].entries[g] = inserted.first->second;
            }
          }
        }
        return 0;
      });
    }
  }

  // 4. Aggregation (aggregate_hash.cpp:1015-1177): chunks outer, aggregates inner, rows innermost; get_or_add_result
  //    (:317-403) with result-id caching in the first key entry.
  const bool has_aggregate_functions = aggregate_count > 0;
  const size_t context_count = has_aggregate_functions ? aggregate_count : 1;
  std::vector<AggregateContext> contexts(context_count);
  for (auto& context : contexts) context.results.reserve(expected_result_size);
  const bool cache_result_ids = context_count > 1 || use_immediate_key_shortcut;

  const auto get_or_add_result = [&](AggregateContext& context, AggregateKey& key, hyb_row_id row_id,
                                     bool cache) -> AggregateResultEntry& {
    auto& results = context.results;
    if (groupby_count == 0) {
      if (results.empty()) {
        results.emplace_back();
        results[0].row_id = row_id;
      }
      return results[0];
    }
    AggregateKeyEntry& first_key_entry = key.entries[0];
    if (cache && (first_key_entry & CACHE_MASK)) {
      const size_t result_id = first_key_entry ^ CACHE_MASK;
      if (result_id >= results.size()) results.resize(static_cast<size_t>(static_cast<double>(result_id + 1) * 1.5));
      results[result_id].row_id = row_id;
      return results[result_id];
    }
    const auto it = context.result_ids.find(key);
    if (it != context.result_ids.end()) {
      const size_t result_id = it->second;
      if (cache) first_key_entry = CACHE_MASK | result_id;
      return results[result_id];
    }
    const size_t result_id = results.size();
    context.result_ids.emplace(key, result_id);
    results.emplace_back();
    results[result_id].row_id = row_id;
    if (cache) first_key_entry = CACHE_MASK | result_id;
    return results[result_id];
  };

  // The Projection below the aggregate (operators/projection.cpp:60-190: one ExpressionEvaluator job per chunk) hands
  // AggregateHash materialised argument columns; the aggregation itself runs chunk after chunk on one thread.
  // arguments[a][c][i] = value of aggregate a's argument for row i of input chunk c.
  std::vector<std::vector<std::vector<Scalar>>> arguments(aggregate_count);
  for (uint32_t a = 0; a < aggregate_count; ++a) {
    if (query->aggregates[a].function == HYB_AGG_COUNT_STAR) continue;
    arguments[a].resize(input_chunks.size());
  }
  run_jobs(input_chunks.size(), threads, [&](size_t c) {
    const auto& rows = *input_chunks[c];
    for (uint32_t a = 0; a < aggregate_count; ++a) {
      const auto& def = query->aggregates[a];
      if (def.function == HYB_AGG_COUNT_STAR) continue;
      auto& column = arguments[a][c];
      column.resize(rows.size());
      for (size_t i = 0; i < rows.size(); ++i) column[i] = evaluate_expression(table, def, rows[i]);
    }
  });

  for (size_t c = 0; c < input_chunks.size(); ++c) {
    const auto& rows = *input_chunks[c];
    auto& keys = keys_per_chunk[c];
    if (!has_aggregate_functions) {
      for (size_t i = 0; i < rows.size(); ++i) get_or_add_result(contexts[0], keys[i], rows[i], use_immediate_key_shortcut);
      continue;
    }
    for (uint32_t a = 0; a < aggregate_count; ++a) {
      const auto& def = query->aggregates[a];
      auto& context = contexts[a];
      if (def.function == HYB_AGG_COUNT_STAR) {
        if (groupby_count == 0) {
          context.results.resize(1);
          context.results[0].aggregate_count += rows.size();
          context.results[0].row_id = hyb_row_id{0, 0};
        } else {
          for (size_t i = 0; i < rows.size(); ++i) {
            ++get_or_add_result(context, keys[i], rows[i], cache_result_ids).aggregate_count;
          }
        }
        continue;
      }
      const bool integral = input_types[a] == HYB_TYPE_INT32 || input_types[a] == HYB_TYPE_INT64;
      for (size_t i = 0; i < rows.size(); ++i) {  // _aggregate_segment (:605-655)
        auto& result = get_or_add_result(context, keys[i], rows[i], cache_result_ids);
        const Scalar value = arguments[a][c][i];
        if (value.is_null) continue;
        switch (def.function) {  // WindowFunctionBuilder (abstract_aggregate_operator.hpp:30-133)
          case HYB_AGG_MIN:
            if (result.aggregate_count == 0 || scalar_less(value, result.extreme)) result.extreme = value;
            break;
          case HYB_AGG_MAX:
            if (result.aggregate_count == 0 || scalar_less(result.extreme, value)) result.extreme = value;
            break;
          case HYB_AGG_SUM:
            if (integral) {
              result.acc_i += scalar_as<int64_t>(value);
            } else {
              result.acc_f += scalar_as<double>(value);
            }
            break;
          case HYB_AGG_AVG:
            result.acc_f += scalar_as<double>(value);
            break;
          default:
            break;
        }
        ++result.aggregate_count;
      }
    }
  }

  // 5. Output (write_groupby_output :421-537, write_aggregate_values :56-176): skip NULL_ROW_ID gaps.
  const auto& first_results = contexts[0].results;
  std::vector<size_t> live;
  for (size_t r = 0; r < first_results.size(); ++r) {
    if (!row_id_is_null(first_results[r].row_id)) live.push_back(r);
  }
  bool synthesize_empty_row = false;
  if (groupby_count == 0 && has_aggregate_functions && live.empty()) synthesize_empty_row = true;  // (:1395-1405)
  const size_t group_count = synthesize_empty_row ? 1 : live.size();
  out->group_count = group_count;
  out->used_immediate_keys = use_immediate_key_shortcut ? 1 : 0;
  out->aggregate_count = aggregate_count;
  out->row_ids = static_cast<hyb_row_id*>(std::malloc(sizeof(hyb_row_id) * std::max<size_t>(group_count, 1)));
  for (size_t g = 0; g < live.size(); ++g) out->row_ids[g] = first_results[live[g]].row_id;
  if (synthesize_empty_row) out->row_ids[0] = NULL_ROW_ID;
  out->columns = static_cast<orc_aggregate_column*>(std::calloc(std::max<uint32_t>(aggregate_count, 1), sizeof(orc_aggregate_column)));
  for (uint32_t a = 0; a < aggregate_count; ++a) {
    const auto& def = query->aggregates[a];
    auto& column = out->columns[a];
    column.value_type = result_types[a];
    const size_t element = (result_types[a] == HYB_TYPE_INT32 || result_types[a] == HYB_TYPE_FLOAT32) ? 4 : 8;
    column.values = std::calloc(std::max<size_t>(group_count, 1), element);
    column.nulls = static_cast<uint8_t*>(std::calloc(std::max<size_t>(group_count, 1), 1));
    const auto& results = contexts[a].results;
    const bool integral = input_types[a] == HYB_TYPE_INT32 || input_types[a] == HYB_TYPE_INT64;
    for (size_t g = 0; g < group_count; ++g) {
      AggregateResultEntry entry;
      if (!synthesize_empty_row && live[g] < results.size()) entry = results[live[g]];
      const bool is_count = def.function == HYB_AGG_COUNT || def.function == HYB_AGG_COUNT_STAR;
      if (is_count) {
        static_cast<int64_t*>(column.values)[g] = static_cast<int64_t>(entry.aggregate_count);
        continue;
      }
      if (entry.aggregate_count == 0) {
        column.nulls[g] = 1;
        continue;
      }
      switch (def.function) {
        case HYB_AGG_SUM:
          if (integral) {
            static_cast<int64_t*>(column.values)[g] = entry.acc_i;
          } else {
            static_cast<double*>(column.values)[g] = entry.acc_f;
          }
          break;
        case HYB_AGG_AVG:
          static_cast<double*>(column.values)[g] = entry.acc_f / static_cast<double>(entry.aggregate_count);
          break;
        default:  // MIN / MAX
          with_type(result_types[a], [&](auto tag) {
            using T = decltype(tag);
            static_cast<T*>(column.values)[g] = scalar_as<T>(entry.extreme);
            return 0;
          });
          break;
      }
    }
  }
  return HYB_OK;
}

void orc_aggregate_result_free(orc_aggregate_result* result) {
  if (!result) return;
  std::free(result->row_ids);
  if (result->columns) {
    for (uint32_t a = 0; a < result->aggregate_count; ++a) {
      std::free(result->columns[a].values);
      std::free(result->columns[a].nulls);
    }
    std::free(result->columns);
  }
  std::memset(result, 0, sizeof(*result));
}

}  // extern "C"
