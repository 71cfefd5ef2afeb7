// The three operators of the hot path driven from C++ through the host-side mirror (include/hyrise_b200.hpp) — the same
// plan bench.py times: TableScan l_shipdate < '1995-01-01'; JoinHash orders x lineitem on orderkey; AggregateHash Q1 with
// its predicate and Projection arithmetic fused. Needs a B200; build with `make example`, run `build/tpch_operators 1`.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "hyrise_b200.hpp"
#include "hyrise_b200_tpch.h"

using namespace hyrise_b200;  // NOLINT

namespace {

ScanPredicate date_predicate(const hyb_tpch* tables, const hyb_table_view& view, ColumnID column, PredicateCondition condition,
                             int32_t year, int32_t month, int32_t day) {
  // the operator shim's job for a string column: per-chunk DictionarySegment::lower_bound / upper_bound as value-IDs
  ScanPredicate predicate{column, condition, std::nullopt, std::nullopt, {}};
  const int32_t day_number = hyb_tpch_day_number(year, month, day);
  predicate.string_value_id_bounds.resize(size_t{view.chunk_count} * 2);
  if (hyb_tpch_value_id_bounds(tables, 0, column, &day_number, 1, predicate.string_value_id_bounds.data()) != 0) {
    throw std::logic_error("hyb_tpch_value_id_bounds failed");
  }
  return predicate;
}

hyb_expr_node column_node(ColumnID column) {
  hyb_expr_node node{};
  node.op = HYB_EXPR_COLUMN;
  node.column_id = column;
  return node;
}

hyb_expr_node literal_one() {
  hyb_expr_node node{};
  node.op = HYB_EXPR_LITERAL;
  node.literal_type = HYB_TYPE_INT32;
  node.literal.i32 = 1;
  return node;
}

hyb_expr_node op_node(hyb_expr_op op) {
  hyb_expr_node node{};
  node.op = op;
  return node;
}

}  // namespace

int main(int argc, char** argv) {
  const double scale_factor = argc > 1 ? std::atof(argv[1]) : 0.1;
  hyb_tpch* tables = nullptr;
  if (hyb_tpch_generate(scale_factor, 42, 0, nullptr, nullptr, &tables) != 0) {
    std::fprintf(stderr, "generator failed\n");
    return 1;
  }
  hyb_table_view lineitem_view{}, orders_view{};
  uint64_t lineitem_rows = 0, orders_rows = 0;
  hyb_tpch_lineitem(tables, &lineitem_view, &lineitem_rows);
  hyb_tpch_orders(tables, &orders_view, &orders_rows);

  try {
    const auto context = std::make_shared<DeviceContext>(0);
    const auto lineitem = std::make_shared<const DeviceTable>(context, lineitem_view);
    const auto orders = std::make_shared<const DeviceTable>(context, orders_view);

    TableScan scan(lineitem, date_predicate(tables, lineitem_view, HYB_L_SHIPDATE, PredicateCondition::LessThan, 1995, 1, 1));
    scan.execute();
    std::printf("TableScan      %llu of %llu rows match, %.3f ms on device\n",
                static_cast<unsigned long long>(scan.get_output()->size()), static_cast<unsigned long long>(lineitem_rows),
                scan.performance_data.device_ms);

    JoinHash join(orders, lineitem, JoinMode::Inner, OperatorJoinPredicate{{HYB_O_ORDERKEY, HYB_L_ORDERKEY}});
    join.execute();
    std::printf("JoinHash       %llu pairs, radix_bits %d, %.3f ms on device\n",
                static_cast<unsigned long long>(join.get_output()->size()), join.get_output()->radix_bits(),
                join.performance_data.device_ms);

    // the scan's output as a reference-table input of a second join: only lineitems shipped before 1995
    JoinHash filtered_join(orders, scan.get_output(), JoinMode::Semi, OperatorJoinPredicate{{HYB_O_ORDERKEY, HYB_L_ORDERKEY}});
    filtered_join.execute();
    std::printf("JoinHash(Semi) %llu lineitems of the scan output have an order\n",
                static_cast<unsigned long long>(filtered_join.get_output()->size()));

    const std::vector<hyb_expr_node> disc_price = {column_node(HYB_L_EXTENDEDPRICE), literal_one(), column_node(HYB_L_DISCOUNT),
                                                   op_node(HYB_EXPR_SUB), op_node(HYB_EXPR_MUL)};
    std::vector<hyb_expr_node> charge = disc_price;
    charge.push_back(literal_one());
    charge.push_back(column_node(HYB_L_TAX));
    charge.push_back(op_node(HYB_EXPR_ADD));
    charge.push_back(op_node(HYB_EXPR_MUL));
    const std::vector<AggregateDefinition> q1 = {
        AggregateDefinition::on_column(WindowFunction::Sum, HYB_L_QUANTITY),
        AggregateDefinition::on_column(WindowFunction::Sum, HYB_L_EXTENDEDPRICE),
        AggregateDefinition{WindowFunction::Sum, disc_price},
        AggregateDefinition{WindowFunction::Sum, charge},
        AggregateDefinition::on_column(WindowFunction::Avg, HYB_L_QUANTITY),
        AggregateDefinition::on_column(WindowFunction::Avg, HYB_L_EXTENDEDPRICE),
        AggregateDefinition::on_column(WindowFunction::Avg, HYB_L_DISCOUNT),
        AggregateDefinition::count_star()};
    AggregateHash aggregate(lineitem, q1, {HYB_L_RETURNFLAG, HYB_L_LINESTATUS},
                            {date_predicate(tables, lineitem_view, HYB_L_SHIPDATE, PredicateCondition::LessThanEquals, 1998, 9, 2)});
    aggregate.execute();
    const auto& groups = aggregate.get_output();
    std::printf("AggregateHash  %llu groups, %.3f ms on device\n", static_cast<unsigned long long>(groups->group_count()),
                aggregate.performance_data.device_ms);
    const auto sum_quantity = groups->values(0).first;
    const auto count_order = groups->values(7).first;
    const auto representatives = groups->group_row_ids();
    for (size_t g = 0; g < sum_quantity.size(); ++g) {
      std::printf("  group %zu (row %u:%u)  sum_qty %.1f  count %.0f\n", g, representatives[g].chunk_id,
                  representatives[g].chunk_offset, sum_quantity[g], count_order[g]);
    }
  } catch (const UnsupportedOnDevice& error) {
    std::fprintf(stderr, "not on the device path: %s\n", error.what());
    hyb_tpch_free(tables);
    return 2;
  } catch (const std::exception& error) {
    std::fprintf(stderr, "failed: %s\n", error.what());
    hyb_tpch_free(tables);
    return 1;
  }
  hyb_tpch_free(tables);
  return 0;
}
