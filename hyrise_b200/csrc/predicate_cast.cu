// Predicate normalisation, host side: what TableScan::create_impl does to a literal before any scan implementation sees
// it (src/lib/operators/table_scan.cpp:340-366 for `column <op> value` / `value <op> column`, :399-441 for BETWEEN) —
// lossless_predicate_variant_cast (src/lib/utils/lossless_predicate_cast.hpp:20-66, .cpp:14-73) on top of lossless_cast
// (src/lib/lossless_cast.hpp:31-176), flip_predicate_condition / between_to_conditions / conditions_to_between
// (src/lib/types.cpp:51-153). The kernels compare in the COLUMN's type, so `float_col < 3.1` must arrive as
// `float_col <= 3.0999999` and `int_col = 16.25` must not arrive at all (HYB_ERR_UNSUPPORTED: the reference falls back to
// the ExpressionEvaluator; the shim runs the CPU operator). No device code in this file.
#include <cmath>
#include <cstdint>
#include <limits>

#include "internal.hpp"

namespace hyb {
namespace {

constexpr double kMaxFloatAsDouble = 340282346638528859811704183484516925440.0;  // largest double a float can hold

struct Number {  // a literal of one of the four numeric types
  int32_t type;
  hyb_value value;
};

// lossless_cast<Target>(Source): `out` receives the value in `target` type; false when information would be lost.
bool lossless_number_cast(const Number& in, int32_t target, hyb_value* out) {
  *out = hyb_value{};
  switch (in.type) {
    case HYB_TYPE_INT32:
      switch (target) {
        case HYB_TYPE_INT32:
          out->i32 = in.value.i32;
          return true;
        case HYB_TYPE_INT64:
          out->i64 = in.value.i32;
          return true;
        case HYB_TYPE_FLOAT32: {  // integral -> floating point: must survive the round trip (lossless_cast.hpp:101-110)
          const float f = static_cast<float>(in.value.i32);
          // the comparison in double avoids the undefined float -> int32 conversion of 2^31
          if (static_cast<double>(f) != static_cast<double>(in.value.i32)) return false;
          out->f32 = f;
          return true;
        }
        default:
          out->f64 = static_cast<double>(in.value.i32);
          return true;
      }
    case HYB_TYPE_INT64:
      switch (target) {
        case HYB_TYPE_INT32:
          if (in.value.i64 < std::numeric_limits<int32_t>::min() || in.value.i64 > std::numeric_limits<int32_t>::max()) return false;
          out->i32 = static_cast<int32_t>(in.value.i64);
          return true;
        case HYB_TYPE_INT64:
          out->i64 = in.value.i64;
          return true;
        case HYB_TYPE_FLOAT32: {
          const float f = static_cast<float>(in.value.i64);
          if (f >= 9223372036854775808.0f || static_cast<int64_t>(f) != in.value.i64) return false;
          out->f32 = f;
          return true;
        }
        default: {
          const double d = static_cast<double>(in.value.i64);
          if (d >= 9223372036854775808.0 || static_cast<int64_t>(d) != in.value.i64) return false;
          out->f64 = d;
          return true;
        }
      }
    case HYB_TYPE_FLOAT32:
    case HYB_TYPE_FLOAT64: {
      const bool from_float = in.type == HYB_TYPE_FLOAT32;
      const double source = from_float ? static_cast<double>(in.value.f32) : in.value.f64;
      if (target == HYB_TYPE_INT32 || target == HYB_TYPE_INT64) {
        // floating point -> integral (lossless_cast.hpp:113-147): no fractional part, inside the explicit boundary values
        double integral_part = 0.0;
        if (std::modf(source, &integral_part) != 0.0 || !std::isfinite(source)) return false;
        if (target == HYB_TYPE_INT32) {
          if (from_float ? (in.value.f32 >= 2147483648.0f || in.value.f32 <= -2147483904.0f)
                         : (source >= 2147483648.0 || source <= -2147483649.0)) {
            return false;
          }
          out->i32 = static_cast<int32_t>(source);
        } else {
          if (from_float ? (in.value.f32 >= 9223372036854775808.0f || in.value.f32 <= -9223373136366403584.0f)
                         : (source >= 9223372036854775808.0 || source <= -9223372036854777856.0)) {
            return false;
          }
          out->i64 = static_cast<int64_t>(source);
        }
        return true;
      }
      if (target == HYB_TYPE_FLOAT64) {
        out->f64 = source;
        return true;
      }
      if (from_float) {
        out->f32 = in.value.f32;
        return true;
      }
      // double -> float (lossless_cast.hpp:156-170)
      if (source > kMaxFloatAsDouble || source < -kMaxFloatAsDouble) return false;
      const float casted = static_cast<float>(source);
      if (static_cast<double>(casted) != source) return false;
      out->f32 = casted;
      return true;
    }
    default:
      return false;
  }
}

// next_float_towards (lossless_predicate_cast.cpp:14-38)
bool next_float_towards(double value, double towards, float* out) {
  if (value > kMaxFloatAsDouble || value < -kMaxFloatAsDouble) return false;
  if (value == towards) return false;
  const float casted = static_cast<float>(value);
  if ((static_cast<double>(casted) < value && towards < value) || (static_cast<double>(casted) > value && towards > value)) {
    *out = casted;
    return true;
  }
  const float next = std::nexttowardf(casted, static_cast<long double>(towards));
  if (!std::isfinite(next)) return false;
  *out = next;
  return true;
}

bool is_binary_numeric_condition(int32_t condition) {  // types.cpp: =, !=, <, <=, >, >=
  return condition >= HYB_PRED_EQUALS && condition <= HYB_PRED_GREATER_THAN_EQUALS;
}

// lossless_predicate_cast<Output>(condition, input) (lossless_predicate_cast.hpp:20-62)
bool predicate_cast(int32_t condition, const Number& literal, int32_t column_type, int32_t* out_condition, hyb_value* out_value) {
  if (lossless_number_cast(literal, column_type, out_value)) {
    *out_condition = condition;
    return true;
  }
  if (!is_binary_numeric_condition(condition)) return false;
  if (literal.type == HYB_TYPE_FLOAT64 && column_type == HYB_TYPE_FLOAT32) {
    if (condition == HYB_PRED_EQUALS) return false;
    float adjusted = 0.f;
    if (condition == HYB_PRED_LESS_THAN || condition == HYB_PRED_LESS_THAN_EQUALS) {
      if (!next_float_towards(literal.value.f64, std::numeric_limits<double>::lowest(), &adjusted)) return false;
      *out_condition = HYB_PRED_LESS_THAN_EQUALS;
      out_value->f32 = adjusted;
      return true;
    }
    if (condition == HYB_PRED_GREATER_THAN || condition == HYB_PRED_GREATER_THAN_EQUALS) {
      if (!next_float_towards(literal.value.f64, std::numeric_limits<double>::max(), &adjusted)) return false;
      *out_condition = HYB_PRED_GREATER_THAN_EQUALS;
      out_value->f32 = adjusted;
      return true;
    }
  }
  return false;
}

bool numeric_type(int32_t type) { return type >= HYB_TYPE_INT32 && type <= HYB_TYPE_FLOAT64; }

}  // namespace
}  // namespace hyb

using namespace hyb;

extern "C" {

int hyb_flip_predicate_condition(int32_t condition, int32_t* out_condition) {
  HYB_CHECK(out_condition, HYB_ERR_INVALID, "NULL argument");
  switch (condition) {  // types.cpp:51-82
    case HYB_PRED_EQUALS:
    case HYB_PRED_NOT_EQUALS:
      *out_condition = condition;
      return HYB_OK;
    case HYB_PRED_LESS_THAN:
      *out_condition = HYB_PRED_GREATER_THAN;
      return HYB_OK;
    case HYB_PRED_LESS_THAN_EQUALS:
      *out_condition = HYB_PRED_GREATER_THAN_EQUALS;
      return HYB_OK;
    case HYB_PRED_GREATER_THAN:
      *out_condition = HYB_PRED_LESS_THAN;
      return HYB_OK;
    case HYB_PRED_GREATER_THAN_EQUALS:
      *out_condition = HYB_PRED_LESS_THAN_EQUALS;
      return HYB_OK;
    default:
      return fail(HYB_ERR_INVALID, "Can't flip PredicateCondition " + std::to_string(condition));  // Fail() in the reference
  }
}

int hyb_next_float_towards(double value, double towards, float* out_value, int32_t* out_possible) {
  HYB_CHECK(out_value && out_possible, HYB_ERR_INVALID, "NULL argument");
  *out_possible = next_float_towards(value, towards, out_value) ? 1 : 0;
  return HYB_OK;
}

int hyb_lossless_predicate_cast(int32_t condition, const hyb_literal* literal, int32_t column_type, int32_t value_on_left,
                                int32_t* out_condition, hyb_value* out_value) {
  HYB_CHECK(literal && out_condition && out_value, HYB_ERR_INVALID, "NULL argument");
  HYB_CHECK(numeric_type(literal->data_type) && numeric_type(column_type), HYB_ERR_UNSUPPORTED,
            "only numeric literals against numeric columns are cast here (string predicates travel as value-ID bounds)");
  int32_t working = condition;
  if (value_on_left) {
    // `value <op> column`: the cast treats its input as the right-hand side, so the condition is flipped before and after
    // (table_scan.cpp:340-354); the returned condition is the one of `column <op'> value`, flipped once more by the caller
    // of ColumnVsValueTableScanImpl (:388-390) — both flips are folded in here.
    HYB_TRY(hyb_flip_predicate_condition(condition, &working));
  }
  const Number number{literal->data_type, literal->value};
  if (!predicate_cast(working, number, column_type, out_condition, out_value)) {
    return fail(HYB_ERR_UNSUPPORTED, "the literal has no lossless form in the column's type: ExpressionEvaluator fallback "
                                     "(table_scan.cpp:355-358), i.e. the CPU operator");
  }
  return HYB_OK;
}

int hyb_lossless_between_cast(int32_t condition, const hyb_literal* lower, const hyb_literal* upper, int32_t column_type,
                              int32_t* out_condition, hyb_value* out_lower, hyb_value* out_upper) {
  HYB_CHECK(lower && upper && out_condition && out_lower && out_upper, HYB_ERR_INVALID, "NULL argument");
  HYB_CHECK(condition >= HYB_PRED_BETWEEN_INCLUSIVE && condition <= HYB_PRED_BETWEEN_EXCLUSIVE, HYB_ERR_INVALID,
            "Input was not a between condition.");
  HYB_CHECK(numeric_type(lower->data_type) && numeric_type(upper->data_type) && numeric_type(column_type), HYB_ERR_UNSUPPORTED,
            "only numeric literals against numeric columns are cast here");
  // between_to_conditions (types.cpp:119-132)
  int32_t lower_condition = (condition == HYB_PRED_BETWEEN_INCLUSIVE || condition == HYB_PRED_BETWEEN_UPPER_EXCLUSIVE)
                                ? HYB_PRED_GREATER_THAN_EQUALS
                                : HYB_PRED_GREATER_THAN;
  int32_t upper_condition = (condition == HYB_PRED_BETWEEN_INCLUSIVE || condition == HYB_PRED_BETWEEN_LOWER_EXCLUSIVE)
                                ? HYB_PRED_LESS_THAN_EQUALS
                                : HYB_PRED_LESS_THAN;
  const Number lower_number{lower->data_type, lower->value}, upper_number{upper->data_type, upper->value};
  if (!predicate_cast(lower_condition, lower_number, column_type, &lower_condition, out_lower) ||
      !predicate_cast(upper_condition, upper_number, column_type, &upper_condition, out_upper)) {
    return fail(HYB_ERR_UNSUPPORTED, "a BETWEEN bound has no lossless form in the column's type: ExpressionEvaluator fallback "
                                     "(table_scan.cpp:410-441), i.e. the CPU operator");
  }
  // conditions_to_between (types.cpp:134-153)
  if (lower_condition == HYB_PRED_GREATER_THAN) {
    *out_condition = upper_condition == HYB_PRED_LESS_THAN ? HYB_PRED_BETWEEN_EXCLUSIVE : HYB_PRED_BETWEEN_LOWER_EXCLUSIVE;
  } else {
    *out_condition = upper_condition == HYB_PRED_LESS_THAN ? HYB_PRED_BETWEEN_UPPER_EXCLUSIVE : HYB_PRED_BETWEEN_INCLUSIVE;
  }
  return HYB_OK;
}

}  // extern "C"
