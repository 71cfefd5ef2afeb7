#include "gpu_operators.hpp"

#include <atomic>

#include "device_column_pool.hpp"
#include "expression/expression_utils.hpp"
#include "expression/pqp_column_expression.hpp"
#include "expression/window_function_expression.hpp"
#include "hyrise_b200.h"
#include "operators/join_helper/join_output_writing.hpp"
#include "resolve_type.hpp"
#include "storage/dictionary_segment.hpp"
#include "storage/pos_lists/row_id_pos_list.hpp"
#include "storage/reference_segment.hpp"
#include "storage/value_segment.hpp"
#include "utils/assert.hpp"
#include "utils/lossless_predicate_cast.hpp"

namespace hyrise {

namespace {

std::atomic<DeviceColumnPool*> global_pool{nullptr};

static_assert(sizeof(RowID) == sizeof(hyb_row_id), "RowID {ChunkID, ChunkOffset} and hyb_row_id share their layout (types.hpp:97-117)");

// DataType is {Null, Int, Long, Float, Double, String} (all_type_variant.hpp:52); hyb_data_type starts at Int.
int32_t hyb_type(DataType data_type) {
  return static_cast<int32_t>(data_type) - 1;
}

hyb_value to_hyb_value(const AllTypeVariant& variant, DataType column_type) {
  auto value = hyb_value{};
  switch (column_type) {
    case DataType::Int:
      value.i32 = boost::get<int32_t>(variant);
      break;
    case DataType::Long:
      value.i64 = boost::get<int64_t>(variant);
      break;
    case DataType::Float:
      value.f32 = boost::get<float>(variant);
      break;
    case DataType::Double:
      value.f64 = boost::get<double>(variant);
      break;
    default:
      Fail("string literals travel as value-ID bounds");
  }
  return value;
}

// DictionarySegment::lower_bound / upper_bound per chunk (dictionary_segment.cpp:94-119): what the reference's dictionary scan
// does at the start of every chunk (column_vs_value_table_scan_impl.cpp:63-139), here once per call for all chunks.
bool string_value_id_bounds(const Table& stored, ColumnID column_id, const AllTypeVariant& value,
                            const std::optional<AllTypeVariant>& value2, std::vector<uint32_t>& bounds) {
  const auto chunk_count = stored.chunk_count();
  const auto values = value2 ? 2u : 1u;
  bounds.resize(size_t{chunk_count} * values * 2);
  for (auto chunk_id = ChunkID{0}; chunk_id < chunk_count; ++chunk_id) {
    const auto segment = std::dynamic_pointer_cast<const BaseDictionarySegment>(stored.get_chunk(chunk_id)->get_segment(column_id));
    if (!segment) {
      return false;
    }
    auto* out = bounds.data() + size_t{chunk_id} * values * 2;
    out[0] = segment->lower_bound(value);
    out[1] = segment->upper_bound(value);
    if (value2) {
      out[2] = segment->lower_bound(*value2);
      out[3] = segment->upper_bound(*value2);
    }
  }
  return true;
}

// PosLists of a device result -> one output chunk of ReferenceSegments per input chunk with matches (table_scan.cpp:199-217).
std::shared_ptr<const Table> wrap_pos_lists(hyb_context* context, const std::shared_ptr<const Table>& stored, hyb_pos_list_t result) {
  auto chunk_count = uint32_t{0};
  auto total = uint64_t{0};
  Assert(hyb_pos_list_info(context, result, &total, &chunk_count) == HYB_OK, hyb_last_error());
  auto offsets = std::vector<uint64_t>(size_t{chunk_count} + 1);
  Assert(hyb_pos_list_chunk_offsets(context, result, offsets.data()) == HYB_OK, hyb_last_error());
  auto output_chunks = std::vector<std::shared_ptr<Chunk>>{};
  for (auto chunk_id = ChunkID{0}; chunk_id < chunk_count; ++chunk_id) {
    const auto matches = offsets[chunk_id + 1] - offsets[chunk_id];
    if (matches == 0) {
      continue;
    }
    auto pos_list = std::make_shared<RowIDPosList>(matches);
    Assert(hyb_pos_list_copy(context, result, offsets[chunk_id], matches, reinterpret_cast<hyb_row_id*>(pos_list->data())) == HYB_OK,
           hyb_last_error());
    pos_list->guarantee_single_chunk();
    auto segments = Segments{};
    for (auto column_id = ColumnID{0}; column_id < stored->column_count(); ++column_id) {
      segments.emplace_back(std::make_shared<ReferenceSegment>(stored, column_id, pos_list));
    }
    output_chunks.emplace_back(std::make_shared<Chunk>(std::move(segments)));
  }
  hyb_pos_list_free(context, result);
  return std::make_shared<Table>(stored->column_definitions(), TableType::References, std::move(output_chunks));
}

}  // namespace

DeviceColumnPool* gpu_column_pool() {
  return global_pool.load();
}

void set_gpu_column_pool(DeviceColumnPool* pool) {
  global_pool.store(pool);
}

std::shared_ptr<const Table> try_gpu_table_scan(const std::shared_ptr<const Table>& in_table, ColumnID column_id,
                                                PredicateCondition condition, const AllTypeVariant& value,
                                                const std::optional<AllTypeVariant>& value2) {
  auto* pool = gpu_column_pool();
  if (!pool || in_table->type() != TableType::Data) {
    return nullptr;  // reference-table inputs: the shim keeps the previous operator's hyb_pos_list_t next to the table it
                     // produced and passes it as input_filter; omitted here for brevity
  }
  const auto* entry = pool->find(in_table);
  if (!entry) {
    return nullptr;
  }
  const auto column_type = in_table->column_data_type(column_id);
  auto predicate = hyb_scan_predicate{};
  predicate.column_id = column_id;
  predicate.condition = static_cast<int32_t>(condition);  // PredicateCondition and hyb_predicate_condition share values
  auto bounds = std::vector<uint32_t>{};
  const auto needs_value = condition != PredicateCondition::IsNull && condition != PredicateCondition::IsNotNull;
  if (needs_value && column_type == DataType::String) {
    if (!string_value_id_bounds(*in_table, column_id, value, value2, bounds)) {
      return nullptr;
    }
    predicate.value_id_bounds = bounds.data();
  } else if (needs_value) {
    // table_scan.cpp:340-366 / :399-441: literals of another type are cast losslessly (or the scan runs on the CPU)
    if (is_between_predicate_condition(condition)) {
      const auto lower = hyb_literal{hyb_type(data_type_from_all_type_variant(value)),
                                     to_hyb_value(value, data_type_from_all_type_variant(value))};
      const auto upper = hyb_literal{hyb_type(data_type_from_all_type_variant(*value2)),
                                     to_hyb_value(*value2, data_type_from_all_type_variant(*value2))};
      if (hyb_lossless_between_cast(predicate.condition, &lower, &upper, hyb_type(column_type), &predicate.condition,
                                    &predicate.lower, &predicate.upper) != HYB_OK) {
        return nullptr;
      }
    } else {
      const auto literal = hyb_literal{hyb_type(data_type_from_all_type_variant(value)),
                                       to_hyb_value(value, data_type_from_all_type_variant(value))};
      if (hyb_lossless_predicate_cast(predicate.condition, &literal, hyb_type(column_type), 0, &predicate.condition,
                                      &predicate.lower) != HYB_OK) {
        return nullptr;
      }
    }
  }
  auto result = hyb_pos_list_t{};
  const auto status = hyb_table_scan(pool->context(), entry->handle, &predicate, 0, &result);
  if (status == HYB_ERR_UNSUPPORTED) {
    return nullptr;
  }
  Assert(status == HYB_OK, hyb_last_error());
  return wrap_pos_lists(pool->context(), in_table, result);
}

std::shared_ptr<const Table> try_gpu_join_hash(const std::shared_ptr<const Table>& build_table,
                                               const std::shared_ptr<const Table>& probe_table, ColumnID build_column_id,
                                               ColumnID probe_column_id, JoinMode mode, size_t radix_bits,
                                               bool build_is_left_input) {
  auto* pool = gpu_column_pool();
  if (!pool) {
    return nullptr;
  }
  const auto* build_entry = pool->find(build_table);
  const auto* probe_entry = pool->find(probe_table);
  if (!build_entry || !probe_entry || build_table->type() != TableType::Data || probe_table->type() != TableType::Data) {
    return nullptr;
  }
  const auto build = hyb_join_side{build_entry->handle, build_column_id, 0};
  const auto probe = hyb_join_side{probe_entry->handle, probe_column_id, 0};
  auto result = hyb_join_result_t{};
  const auto status = hyb_join_hash(pool->context(), &build, &probe, static_cast<int32_t>(mode), static_cast<int32_t>(radix_bits), &result);
  if (status == HYB_ERR_UNSUPPORTED) {
    return nullptr;  // string / float keys, secondary predicates
  }
  Assert(status == HYB_OK, hyb_last_error());
  // The pairs arrive in the order probe() emits them (partition, probe row, build row): one PosList pair per partition, then
  // the reference's own write_output_chunks (join_helper/join_output_writing.cpp:205-340) merges / splits and builds the chunks.
  auto pair_count = uint64_t{0};
  auto partition_count = uint32_t{0};
  Assert(hyb_join_result_info(pool->context(), result, &pair_count, &partition_count, nullptr) == HYB_OK, hyb_last_error());
  auto offsets = std::vector<uint64_t>(size_t{partition_count} + 1);
  Assert(hyb_join_result_partition_offsets(pool->context(), result, offsets.data()) == HYB_OK, hyb_last_error());
  const auto emits_build_side = mode != JoinMode::Semi && mode != JoinMode::AntiNullAsTrue && mode != JoinMode::AntiNullAsFalse;
  auto build_pos_lists = std::vector<RowIDPosList>(partition_count);
  auto probe_pos_lists = std::vector<RowIDPosList>(partition_count);
  for (auto partition = size_t{0}; partition < partition_count; ++partition) {
    const auto rows = offsets[partition + 1] - offsets[partition];
    probe_pos_lists[partition].resize(rows);
    if (emits_build_side) {
      build_pos_lists[partition].resize(rows);
    }
    Assert(hyb_join_result_copy(pool->context(), result, offsets[partition], rows,
                                emits_build_side ? reinterpret_cast<hyb_row_id*>(build_pos_lists[partition].data()) : nullptr,
                                reinterpret_cast<hyb_row_id*>(probe_pos_lists[partition].data())) == HYB_OK,
           hyb_last_error());
  }
  hyb_join_result_free(pool->context(), result);
  auto& left_pos_lists = build_is_left_input ? build_pos_lists : probe_pos_lists;
  auto& right_pos_lists = build_is_left_input ? probe_pos_lists : build_pos_lists;
  const auto& left_table = build_is_left_input ? build_table : probe_table;
  const auto& right_table = build_is_left_input ? probe_table : build_table;
  auto output_chunks = write_output_chunks(left_pos_lists, right_pos_lists, left_table, right_table, /*create_left_side_pos_lists_by_column=*/false,
                                           /*create_right_side_pos_lists_by_column=*/false, OutputColumnOrder::LeftFirstRightSecond,
                                           /*allow_partition_merge=*/true);
  auto definitions = left_table->column_definitions();
  if (emits_build_side) {
    const auto& right_definitions = right_table->column_definitions();
    definitions.insert(definitions.end(), right_definitions.begin(), right_definitions.end());
  }
  return std::make_shared<Table>(definitions, TableType::References, std::move(output_chunks));
}

std::shared_ptr<const Table> try_gpu_aggregate_hash(const std::shared_ptr<const Table>& in_table,
                                                    const std::vector<ColumnID>& groupby_column_ids,
                                                    const std::vector<std::shared_ptr<WindowFunctionExpression>>& aggregates) {
  auto* pool = gpu_column_pool();
  if (!pool || in_table->type() != TableType::Data) {
    return nullptr;
  }
  const auto* entry = pool->find(in_table);
  if (!entry) {
    return nullptr;
  }
  // function + argument column (the reference has materialised arithmetic in a Projection below the aggregate; a fusing shim
  // passes the Projection's expressions as RPN programs and the scans below it as query.predicates instead)
  auto nodes = std::vector<hyb_expr_node>(aggregates.size());
  auto definitions = std::vector<hyb_aggregate_def>(aggregates.size());
  for (auto index = size_t{0}; index < aggregates.size(); ++index) {
    const auto& aggregate = *aggregates[index];
    auto& definition = definitions[index];
    definition.function = static_cast<int32_t>(aggregate.window_function);  // WindowFunction and hyb_aggregate_function share values
    const auto column = std::dynamic_pointer_cast<const PQPColumnExpression>(aggregate.argument());
    if (aggregate.window_function == WindowFunction::Count && (!column || column->column_id == INVALID_COLUMN_ID)) {
      definition.function = HYB_AGG_COUNT_STAR;
      continue;
    }
    if (!column) {
      return nullptr;
    }
    nodes[index] = hyb_expr_node{};
    nodes[index].op = HYB_EXPR_COLUMN;
    nodes[index].column_id = column->column_id;
    definition.nodes = &nodes[index];
    definition.node_count = 1;
  }
  auto groupby = std::vector<uint32_t>(groupby_column_ids.begin(), groupby_column_ids.end());
  auto query = hyb_aggregate_query{};
  query.table = entry->handle;
  query.groupby_count = static_cast<uint32_t>(groupby.size());
  query.groupby_column_ids = groupby.data();
  query.aggregate_count = static_cast<uint32_t>(definitions.size());
  query.aggregates = definitions.data();
  auto result = hyb_aggregate_result_t{};
  const auto status = hyb_aggregate_hash(pool->context(), &query, &result);
  if (status == HYB_ERR_UNSUPPORTED) {
    return nullptr;  // COUNT DISTINCT, STDDEV_SAMP, string aggregates
  }
  Assert(status == HYB_OK, hyb_last_error());
  // Output: group-by columns as ReferenceSegments over the representative RowIDs (write_groupby_output, aggregate_hash.cpp:
  // 421-537), aggregate columns as ValueSegments (write_aggregate_output, :1374-1456), split into 65 535-row chunks like
  // split_results_chunk_wise (:237-294). hyb_aggregate_result_row_ids / hyb_aggregate_result_values deliver both in group order.
  auto group_count = uint64_t{0};
  Assert(hyb_aggregate_result_info(pool->context(), result, &group_count, nullptr) == HYB_OK, hyb_last_error());
  auto row_ids = std::make_shared<RowIDPosList>(group_count);
  Assert(hyb_aggregate_result_row_ids(pool->context(), result, reinterpret_cast<hyb_row_id*>(row_ids->data())) == HYB_OK,
         hyb_last_error());
  auto output_definitions = TableColumnDefinitions{};
  auto segments = Segments{};
  for (const auto column_id : groupby_column_ids) {
    output_definitions.emplace_back(in_table->column_definitions()[column_id]);
    segments.emplace_back(std::make_shared<ReferenceSegment>(in_table, column_id, row_ids));
  }
  for (auto index = size_t{0}; index < aggregates.size(); ++index) {
    const auto result_type = aggregates[index]->data_type();  // WindowFunctionTraits (window_function_traits.hpp:14-77)
    resolve_data_type(result_type, [&](const auto type) {
      using AggregateType = typename decltype(type)::type;
      if constexpr (!std::is_same_v<AggregateType, pmr_string>) {
        auto values = pmr_vector<AggregateType>(group_count);
        auto null_bytes = std::vector<uint8_t>(group_count);
        auto value_type = int32_t{0};
        Assert(hyb_aggregate_result_values(pool->context(), result, static_cast<uint32_t>(index), values.data(), null_bytes.data(),
                                           &value_type) == HYB_OK,
               hyb_last_error());
        Assert(value_type == hyb_type(result_type), "result type mismatch between the shim and the library");
        auto nulls = pmr_vector<bool>(null_bytes.begin(), null_bytes.end());
        segments.emplace_back(std::make_shared<ValueSegment<AggregateType>>(std::move(values), std::move(nulls)));
      }
    });
    output_definitions.emplace_back(aggregates[index]->as_column_name(), result_type, true);
  }
  hyb_aggregate_result_free(pool->context(), result);
  auto output = std::make_shared<Table>(output_definitions, TableType::Data);
  if (group_count > 0) {
    output->append_chunk(segments);  // > 65 535 groups: slice the vectors per Chunk::DEFAULT_SIZE first (:237-294)
  }
  return output;
}

}  // namespace hyrise
