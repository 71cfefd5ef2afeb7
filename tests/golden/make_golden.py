#!/usr/bin/env python
"""Copies the reference's own golden inputs/expectations for the hot path into tests/golden/ (run in the build
container where /root/reference is mounted; the GPU box has no /root/reference, so tests only read tests/golden/).

Sources (all under /root/reference/resources/test_data/tbl, hyrise/hyrise @ 2f7bedf3):
  * scan fixtures used by src/test/lib/operators/table_scan_test.cpp and table_scan_between_test.cpp
  * join_test_runner/input_table_{left,right}_{0,10,15}.tbl (src/test/lib/operators/join_test_runner.cpp:184-543)
  * int_int4_with_null.tbl / int_float_with_null.tbl (src/test/lib/operators/join_hash/join_hash_steps_test.cpp)
  * aggregateoperator/** (src/test/lib/operators/aggregate_test.cpp:262-850)
  * tpch/sf-0.001 and sf-0.01 lineitem/orders, reduced to the columns on the hot path (join_hash_test.cpp:26-46,
    tpch_queries.cpp Q1/Q3/Q6) and stored as compressed .npz
Data files only — no reference source code is copied.
"""
import os
import shutil
import sys

import numpy as np

REF = "/root/reference/resources/test_data/tbl"
HERE = os.path.dirname(os.path.abspath(__file__))

TBL = [
    "int_float.tbl", "int_float_filtered.tbl", "int_float_filtered2.tbl", "int_sorted.tbl", "int_sorted_filtered.tbl",
    "int_sorted_filtered2.tbl", "int_int_shuffled.tbl", "int_int_shuffled_2.tbl", "int_float_with_null.tbl",
    "int_int_w_null_8_rows.tbl", "int_int4_with_null.tbl", "int_only_null.tbl", "int_empty.tbl",
    "int_empty_nullable.tbl", "int_float_null_1.tbl", "int_float_null_2.tbl", "int_with_nulls_large.tbl",
    "long_with_null.tbl", "int_int_int_null.tbl", "int_string.tbl", "int_float2.tbl", "int_float4.tbl",
    "float_int.tbl", "int_int.tbl", "int_int2.tbl", "int_int3.tbl", "int_float_double_string.tbl",
]

LINEITEM_COLUMNS = ["l_orderkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag",
                    "l_linestatus", "l_shipdate"]
ORDERS_COLUMNS = ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"]
NP_TYPES = {"int": np.int32, "long": np.int64, "float": np.float32, "double": np.float64}


def tbl_to_npz(source, destination, keep):
    with open(source) as handle:
        lines = handle.read().split("\n")
    if lines[-1] == "":
        lines.pop()
    names = lines[0].split("|")
    types = lines[1].split("|")
    index = [names.index(name) for name in keep]
    rows = [line.split("|") for line in lines[2:]]
    arrays = {}
    for name, column in zip(keep, index):
        raw = [row[column] for row in rows]
        base = types[column].split("_")[0]
        arrays[name] = np.array(raw, dtype="S") if base == "string" else np.array(raw, dtype=np.float64).astype(
            NP_TYPES[base]) if base in ("float", "double") else np.array(raw, dtype=NP_TYPES[base])
        arrays["type__" + name] = np.array(base)
    np.savez_compressed(destination, **arrays)


def copy_binary_fixtures():
    """resources/test_data/bin: the reference's own binary-table fixtures (binary_parser_test.cpp), a few hundred bytes each.
    LZ4MultipleBlocks.bin (49 KB) is left out: LZ4 segments are not on the device path."""
    source_root = "/root/reference/resources/test_data/bin"
    target_root = os.path.join(HERE, "bin")
    os.makedirs(target_root, exist_ok=True)
    for entry in sorted(os.listdir(source_root)):
        source = os.path.join(source_root, entry)
        if os.path.isdir(source):
            os.makedirs(os.path.join(target_root, entry), exist_ok=True)
            for name in sorted(os.listdir(source)):
                if name.endswith(".bin"):
                    shutil.copyfile(os.path.join(source, name), os.path.join(target_root, entry, name))
        elif entry.endswith(".bin") and entry != "LZ4MultipleBlocks.bin":
            shutil.copyfile(source, os.path.join(target_root, entry))


def main():
    copy_binary_fixtures()
    if not os.path.isdir(REF):
        sys.exit("reference test data not mounted; nothing to do")
    out_tbl = os.path.join(HERE, "tbl")
    os.makedirs(out_tbl, exist_ok=True)
    for name in TBL:
        shutil.copyfile(os.path.join(REF, name), os.path.join(out_tbl, name))
    for directory in ("join_test_runner", "aggregateoperator"):
        target = os.path.join(out_tbl, directory)
        shutil.rmtree(target, ignore_errors=True)
        shutil.copytree(os.path.join(REF, directory), target)
    for root, _, files in os.walk(out_tbl):
        for name in files:
            os.chmod(os.path.join(root, name), 0o644)
        os.chmod(root, 0o755)
    tpch = os.path.join(HERE, "tpch")
    os.makedirs(tpch, exist_ok=True)
    for scale in ("sf-0.001", "sf-0.01"):
        tbl_to_npz(os.path.join(REF, "tpch", scale, "lineitem.tbl"), os.path.join(tpch, f"{scale}_lineitem.npz"),
                   LINEITEM_COLUMNS)
        tbl_to_npz(os.path.join(REF, "tpch", scale, "orders.tbl"), os.path.join(tpch, f"{scale}_orders.npz"),
                   ORDERS_COLUMNS)
    # Q3 (tpch_queries.cpp:92-109) additionally filters customer by c_mktsegment
    tbl_to_npz(os.path.join(REF, "tpch", "sf-0.01", "customer.tbl"), os.path.join(tpch, "sf-0.01_customer.npz"),
               ["c_custkey", "c_mktsegment"])


if __name__ == "__main__":
    main()
