// Host-only check of the C++ mirror's predicate normalisation (include/hyrise_b200.hpp). The expectations are the ones the
// reference states in src/test/lib/utils/lossless_predicate_cast_test.cpp:49-91 (numeric types; strings never reach the
// device path) and src/test/lib/types_test.cpp for flip_predicate_condition. No GPU needed: these calls are host logic.
#include <cstdio>
#include <cstdlib>

#include "hyrise_b200.hpp"

using namespace hyrise_b200;

static int failures = 0;
#define EXPECT(condition)                                                  \
  do {                                                                     \
    if (!(condition)) {                                                    \
      std::fprintf(stderr, "line %d: %s\n", __LINE__, #condition);         \
      ++failures;                                                          \
    }                                                                      \
  } while (0)

template <typename T>
static bool is(const std::optional<std::pair<PredicateCondition, AllTypeVariant>>& result, PredicateCondition condition, T value) {
  return result && result->first == condition && std::holds_alternative<T>(result->second) && std::get<T>(result->second) == value;
}

int main() {
  using PC = PredicateCondition;
  // NonFloatTypes
  EXPECT(is(lossless_predicate_variant_cast(PC::GreaterThan, int64_t{10}, DataType::Long), PC::GreaterThan, int64_t{10}));
  EXPECT(is(lossless_predicate_variant_cast(PC::Equals, int64_t{10}, DataType::Long), PC::Equals, int64_t{10}));
  EXPECT(is(lossless_predicate_variant_cast(PC::GreaterThan, int64_t{10}, DataType::Int), PC::GreaterThan, int32_t{10}));
  EXPECT(!lossless_predicate_variant_cast(PC::GreaterThan, int64_t{100'000'000'000}, DataType::Int));
  EXPECT(is(lossless_predicate_variant_cast(PC::GreaterThan, int32_t{10}, DataType::Long), PC::GreaterThan, int64_t{10}));
  // FloatTypeWithLosslessCast
  EXPECT(is(lossless_predicate_variant_cast(PC::GreaterThan, 3.0, DataType::Float), PC::GreaterThan, 3.f));
  // FloatTypeWithAdjustedValues
  EXPECT(is(lossless_predicate_variant_cast(PC::LessThan, 3.1, DataType::Float), PC::LessThanEquals, 3.099999904632568359375f));
  EXPECT(is(lossless_predicate_variant_cast(PC::LessThanEquals, 3.1, DataType::Float), PC::LessThanEquals, 3.099999904632568359375f));
  EXPECT(!lossless_predicate_variant_cast(PC::Equals, 3.1, DataType::Float));
  EXPECT(is(lossless_predicate_variant_cast(PC::GreaterThan, 3.1, DataType::Float), PC::GreaterThanEquals, 3.1000001430511474609375f));
  EXPECT(is(lossless_predicate_variant_cast(PC::GreaterThanEquals, 3.1, DataType::Float), PC::GreaterThanEquals,
            3.1000001430511474609375f));

  // flip_predicate_condition (types.cpp:51-82)
  EXPECT(flip_predicate_condition(PC::LessThan) == PC::GreaterThan);
  EXPECT(flip_predicate_condition(PC::GreaterThanEquals) == PC::LessThanEquals);
  EXPECT(flip_predicate_condition(PC::Equals) == PC::Equals);
  bool threw = false;
  try {
    flip_predicate_condition(PC::BetweenInclusive);
  } catch (const std::exception&) {
    threw = true;
  }
  EXPECT(threw);

  // ScanPredicate::normalized — what TableScan::create_impl does with a literal of another type (table_scan.cpp:340-366,
  // :399-441): int column BETWEEN 2.5 AND 7.0 cannot be cast losslessly on the lower bound.
  ScanPredicate less{0, PC::LessThan, AllTypeVariant{3.1}, std::nullopt, {}};
  const auto normalized = less.normalized(DataType::Float);
  EXPECT(normalized && normalized->condition == PC::LessThanEquals && std::get<float>(*normalized->value) == 3.099999904632568359375f);
  ScanPredicate between{0, PC::BetweenInclusive, AllTypeVariant{int64_t{2}}, AllTypeVariant{int64_t{7}}, {}};
  const auto narrowed = between.normalized(DataType::Int);
  EXPECT(narrowed && narrowed->condition == PC::BetweenInclusive && std::get<int32_t>(*narrowed->value) == 2 &&
         std::get<int32_t>(*narrowed->value2) == 7);
  ScanPredicate fractional{0, PC::BetweenInclusive, AllTypeVariant{2.5}, AllTypeVariant{7.0}, {}};
  EXPECT(!fractional.normalized(DataType::Int));

  // BinaryTable: the reference's fixture AllTypesNullValues/Dictionary.bin (binary_parser_test.cpp:165-188) parses on the host
  if (const char* fixture = std::getenv("HYB_BINARY_FIXTURE")) {
    BinaryTable file(fixture, /*pinned=*/false);
    EXPECT(file.column_count() == 5 && file.chunk_count() == 1);
    EXPECT(file.column_name(0) == "a" && file.column_name(3) == "d");
    const auto bounds = file.value_id_bounds(3, "one");   // column d holds {"one","two","three",NULL,"five"}: sorted five, one, three, two
    EXPECT(bounds.size() == 2 && bounds[0] == 1 && bounds[1] == 2);
    bool threw_missing = false;
    try {
      BinaryTable missing("not_existing_file", false);
    } catch (const std::exception&) {
      threw_missing = true;
    }
    EXPECT(threw_missing);
  }

  if (failures) return EXIT_FAILURE;
  std::puts("predicate casts OK");
  return EXIT_SUCCESS;
}
