"""Predicate normalisation (a18): the oracle's restatement and the product's host helper, both pinned to the literal
expectations of the reference's src/test/lib/utils/lossless_predicate_cast_test.cpp (:14-107), then checked against each
other on random literals. Host-only code: no GPU needed."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as orc
from hyrise_b200 import capi

I32, I64, F32, F64 = capi.TYPE_INT32, capi.TYPE_INT64, capi.TYPE_FLOAT32, capi.TYPE_FLOAT64
EQ, NE, LT, LE, GT, GE = (capi.PRED_EQUALS, capi.PRED_NOT_EQUALS, capi.PRED_LESS_THAN, capi.PRED_LESS_THAN_EQUALS,
                          capi.PRED_GREATER_THAN, capi.PRED_GREATER_THAN_EQUALS)
BIG = 340282346638528859811704183484516925440.0      # largest double a float holds
BIGGER = 340282346638528897590636046441678635008.0   # the next double


def value_of(data_type, value):
    v = capi.Value()
    if data_type == I32:
        v.i32 = int(value)
    elif data_type == I64:
        v.i64 = int(value)
    elif data_type == F32:
        v.f32 = float(np.float32(value))
    else:
        v.f64 = float(value)
    return v


def read(data_type, v):
    return {I32: v.i32, I64: v.i64, F32: np.float32(v.f32), F64: v.f64}[data_type]


def oracle_cast(condition, literal_type, literal, column_type, value_on_left=False):
    lib = orc.load()
    lib.orc_normalize_predicate.argtypes = [C.c_int32, C.c_int32, capi.Value, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                            C.POINTER(capi.Value)]
    out_condition, out_value = C.c_int32(), capi.Value()
    ok = lib.orc_normalize_predicate(condition, literal_type, value_of(literal_type, literal), column_type, int(value_on_left),
                                     C.byref(out_condition), C.byref(out_value))
    return (out_condition.value, read(column_type, out_value)) if ok else None


def product_cast(condition, literal_type, literal, column_type, value_on_left=False):
    lib = capi.load_library()
    lit = capi.Literal(literal_type, value_of(literal_type, literal))
    out_condition, out_value = C.c_int32(), capi.Value()
    status = lib.hyb_lossless_predicate_cast(condition, C.byref(lit), column_type, int(value_on_left), C.byref(out_condition),
                                             C.byref(out_value))
    if status == capi.HYB_ERR_UNSUPPORTED:
        return None
    capi.check(status)
    return out_condition.value, read(column_type, out_value)


def next_float(impl, value, towards):
    out = C.c_float()
    if impl == "oracle":
        lib = orc.load()
        lib.orc_next_float_towards.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_float)]
        return np.float32(out.value) if lib.orc_next_float_towards(value, towards, C.byref(out)) else None
    possible = C.c_int32()
    capi.check(capi.load_library().hyb_next_float_towards(value, towards, C.byref(out), C.byref(possible)))
    return np.float32(out.value) if possible.value else None


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_next_float_towards(impl):
    # lossless_predicate_cast_test.cpp:14-50
    assert next_float(impl, 3, 2) == np.float32(2.9999997615814208984375)
    assert next_float(impl, 3, 4) == np.float32(3.0000002384185791015625)
    assert next_float(impl, 3.1, 3) == np.float32(3.099999904632568359375)
    assert next_float(impl, 3.1, 4) == np.float32(3.1000001430511474609375)
    assert next_float(impl, 3.1, 3.1) is None
    assert next_float(impl, BIG, 0) == np.float32(340282326356119256160033759537265639424.0)
    assert next_float(impl, BIG, BIG * 10) is None
    assert next_float(impl, BIGGER, 0) is None
    assert next_float(impl, BIGGER, BIGGER * 10) is None
    assert next_float(impl, -BIG, -10) == np.float32(-340282326356119256160033759537265639424.0)
    assert next_float(impl, -BIG, -BIG * 10) is None
    assert next_float(impl, -BIGGER, 10) is None
    assert next_float(impl, -BIGGER, -BIGGER * 10) is None


@pytest.mark.parametrize("cast", [oracle_cast, product_cast])
def test_reference_expectations(cast):
    # NonFloatTypes (:52-76)
    assert cast(GT, I64, 10, I64) == (GT, 10)
    assert cast(EQ, I64, 10, I64) == (EQ, 10)
    assert cast(GT, I64, 10, I32) == (GT, 10)
    assert cast(GT, I64, 100_000_000_000, I32) is None
    assert cast(GT, I32, 10, I64) == (GT, 10)
    # FloatTypeWithLosslessCast (:78-83)
    assert cast(GT, F64, 3.0, F32) == (GT, np.float32(3.0))
    # FloatTypeWithAdjustedValues (:85-98)
    assert cast(LT, F64, 3.1, F32) == (LE, np.float32(3.099999904632568359375))
    assert cast(LE, F64, 3.1, F32) == (LE, np.float32(3.099999904632568359375))
    assert cast(EQ, F64, 3.1, F32) is None
    assert cast(GT, F64, 3.1, F32) == (GE, np.float32(3.1000001430511474609375))
    assert cast(GE, F64, 3.1, F32) == (GE, np.float32(3.1000001430511474609375))
    # table_scan.cpp:323-326: `int_column = 16.25` must not become `int_column = 16`
    assert cast(EQ, F64, 16.25, I32) is None
    assert cast(EQ, F64, 16.0, I32) == (EQ, 16)
    assert cast(NE, F64, 3.1, F32) is None          # not handled by the reference: only the four inequalities are
    # value on the left: `3.1 > float_col` is `float_col < 3.1` -> `float_col <= 3.0999999`
    assert cast(GT, F64, 3.1, F32, True) == (LE, np.float32(3.099999904632568359375))
    assert cast(LE, F64, 3.1, F32, True) == (GE, np.float32(3.1000001430511474609375))
    assert cast(LT, I32, 7, I64, True) == (GT, 7)
    # lossless_cast.hpp boundary values (:124-146)
    assert cast(EQ, F64, 2147483648.0, I32) is None and cast(EQ, F64, 2147483647.0, I32) == (EQ, 2147483647)
    assert cast(EQ, F64, -2147483649.0, I32) is None and cast(EQ, F64, -2147483648.0, I32) == (EQ, -2147483648)
    assert cast(EQ, I32, 16777217, F32) is None and cast(EQ, I32, 16777216, F32) == (EQ, np.float32(16777216))
    assert cast(EQ, I64, 2 ** 53 + 1, F64) is None and cast(EQ, I64, 2 ** 53, F64) == (EQ, float(2 ** 53))


def test_between_composition():
    # table_scan.cpp:399-441: between_to_conditions -> cast both bounds -> conditions_to_between
    lib, product = orc.load(), capi.load_library()
    lib.orc_normalize_between.argtypes = [C.c_int32, C.c_int32, capi.Value, C.c_int32, capi.Value, C.c_int32,
                                          C.POINTER(C.c_int32), C.POINTER(capi.Value), C.POINTER(capi.Value)]
    cases = [
        (capi.PRED_BETWEEN_INCLUSIVE, (F64, 3.1), (F64, 4.1), F32,
         (capi.PRED_BETWEEN_INCLUSIVE, np.float32(3.1000001430511474609375), np.float32(4.099999904632568359375))),
        (capi.PRED_BETWEEN_EXCLUSIVE, (F64, 3.1), (F64, 4.1), F32,   # x > 3.1 -> x >= next; x < 4.1 -> x <= prev
         (capi.PRED_BETWEEN_INCLUSIVE, np.float32(3.1000001430511474609375), np.float32(4.099999904632568359375))),
        (capi.PRED_BETWEEN_EXCLUSIVE, (F64, 3.0), (F64, 4.0), F32, (capi.PRED_BETWEEN_EXCLUSIVE, np.float32(3), np.float32(4))),
        (capi.PRED_BETWEEN_LOWER_EXCLUSIVE, (I64, 3), (F64, 9.0), I32, (capi.PRED_BETWEEN_LOWER_EXCLUSIVE, 3, 9)),
        (capi.PRED_BETWEEN_UPPER_EXCLUSIVE, (I32, 3), (F64, 9.5), I32, None),
    ]
    for condition, (lower_type, lower), (upper_type, upper), column_type, expected in cases:
        out_condition, out_lower, out_upper = C.c_int32(), capi.Value(), capi.Value()
        ok = lib.orc_normalize_between(condition, lower_type, value_of(lower_type, lower), upper_type, value_of(upper_type, upper),
                                       column_type, C.byref(out_condition), C.byref(out_lower), C.byref(out_upper))
        got = (out_condition.value, read(column_type, out_lower), read(column_type, out_upper)) if ok else None
        assert got == expected, (condition, lower, upper)
        lo, hi = capi.Literal(lower_type, value_of(lower_type, lower)), capi.Literal(upper_type, value_of(upper_type, upper))
        status = product.hyb_lossless_between_cast(condition, C.byref(lo), C.byref(hi), column_type, C.byref(out_condition),
                                                   C.byref(out_lower), C.byref(out_upper))
        got = None if status == capi.HYB_ERR_UNSUPPORTED else (out_condition.value, read(column_type, out_lower),
                                                                read(column_type, out_upper))
        assert got == expected, (condition, lower, upper)


def test_product_matches_oracle_on_random_literals():
    rng = np.random.default_rng(3)
    specials = [0, 1, -1, 2 ** 24, 2 ** 24 + 1, 2 ** 31 - 1, -2 ** 31, 2 ** 31, 2 ** 53, 2 ** 53 + 1, 2 ** 63 - 1, -2 ** 63]
    for _ in range(4000):
        literal_type = int(rng.integers(0, 4))
        column_type = int(rng.integers(0, 4))
        condition = int(rng.integers(EQ, GE + 1))
        if literal_type in (I32, I64):
            literal = int(specials[int(rng.integers(0, len(specials)))]) if rng.random() < 0.5 else int(rng.integers(-10 ** 12, 10 ** 12))
            literal = max(min(literal, 2 ** 31 - 1), -2 ** 31) if literal_type == I32 else max(min(literal, 2 ** 63 - 1), -2 ** 63)
        else:
            literal = float(rng.choice([0.5, 3.1, -3.1, 16.0, 1e10, 2147483648.0, -2147483904.0, 1e300, 9.3e18, float(2 ** 24 + 1)]))
            if rng.random() < 0.5:
                literal = float(rng.normal() * 10.0 ** int(rng.integers(-3, 12)))
            if literal_type == F32:
                literal = float(np.float32(np.clip(literal, -3e38, 3e38)))
        left = bool(rng.integers(0, 2))
        assert product_cast(condition, literal_type, literal, column_type, left) == \
            oracle_cast(condition, literal_type, literal, column_type, left), (condition, literal_type, literal, column_type, left)


def test_flip_predicate_condition():
    lib = capi.load_library()
    out = C.c_int32()
    for condition, expected in ((EQ, EQ), (NE, NE), (LT, GT), (LE, GE), (GT, LT), (GE, LE)):
        capi.check(lib.hyb_flip_predicate_condition(condition, C.byref(out)))
        assert out.value == expected
    assert lib.hyb_flip_predicate_condition(capi.PRED_LIKE, C.byref(out)) == capi.HYB_ERR_INVALID   # Fail("Can't flip ...")
