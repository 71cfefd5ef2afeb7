"""Which part of the Q1 aggregate costs what: the same table, progressively smaller queries (development aid)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from bench import *  # noqa
from hyrise_b200.device import DeviceContext
tables = TpchTables(10.0, seed=42)
device = DeviceContext(0)
lineitem = device.upload(tables.lineitem); device.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
def run(name, groupby, aggregates, predicates, n=8):
    xs = []
    for _ in range(n):
        flush.fill_(1); torch.cuda.synchronize()
        out = device.aggregate_hash(lineitem, groupby, aggregates, predicates=predicates); device.synchronize()
        st = device.last_stats(); xs.append(st.dominant_kernel_ms)
    print(f"{name:58s} {np.median(xs[2:]):.3f} ms  groups {out.group_count} launches {st.kernel_launches} bytes {st.algorithmic_bytes/1e6:.0f} MB", flush=True)
col = Expression.column
_ONE = ("lit", capi.TYPE_INT32, 1)
price_chain2 = Expression([("col", L_EXTENDEDPRICE), _ONE, ("col", L_DISCOUNT), "-", "*"])
price_chain3 = Expression([("col", L_EXTENDEDPRICE), _ONE, ("col", L_DISCOUNT), "-", "*", _ONE, ("col", L_TAX), "+", "*"])
S = capi.AGG_SUM
run("Q1 full", Q1_GROUPBY, Q1_AGGREGATES, Q1_PREDICATES)
run("Q1, predicate matches nothing", Q1_GROUPBY, Q1_AGGREGATES, [Predicate(L_SHIPDATE, capi.PRED_LESS_THAN, "1900-01-01")])
run("Q1 without predicate", Q1_GROUPBY, Q1_AGGREGATES, [])
run("Q1 without group-by", [], Q1_AGGREGATES, Q1_PREDICATES)
run("group-by, COUNT(*) only", Q1_GROUPBY, [Aggregate(capi.AGG_COUNT_STAR)], Q1_PREDICATES)
run("group-by, SUM(quantity)", Q1_GROUPBY, [Aggregate(S, col(L_QUANTITY))], Q1_PREDICATES)
run("group-by, SUM(extendedprice)", Q1_GROUPBY, [Aggregate(S, col(L_EXTENDEDPRICE))], Q1_PREDICATES)
run("group-by, SUM(price*(1-disc))", Q1_GROUPBY, [Aggregate(S, price_chain2)], Q1_PREDICATES)
run("group-by, SUM(price*(1-disc)*(1+tax))", Q1_GROUPBY, [Aggregate(S, price_chain3)], Q1_PREDICATES)
run("no group-by, SUM(quantity)", [], [Aggregate(S, col(L_QUANTITY))], [])
run("no group-by, SUM(extendedprice)", [], [Aggregate(S, col(L_EXTENDEDPRICE))], [])
run("no group-by, COUNT(*) with predicate", [], [Aggregate(capi.AGG_COUNT_STAR)], Q1_PREDICATES)
