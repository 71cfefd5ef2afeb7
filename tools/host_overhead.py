"""Where does the wall time of a bench step go? (development aid)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from bench import *  # noqa
from hyrise_b200.device import DeviceContext
tables = TpchTables(10.0, seed=42)
device = DeviceContext(0)
lineitem = device.upload(tables.lineitem); orders = device.upload(tables.orders); device.synchronize()
def t(f, n=10):
    device.synchronize(); xs=[]
    for _ in range(n):
        a=time.perf_counter(); r=f(); device.synchronize(); xs.append(time.perf_counter()-a)
        if hasattr(r,'free'): r.free()
    return round(float(np.median(xs))*1e3,3)
print("scan      ", t(lambda: device.table_scan(lineitem, SCAN_PREDICATE)))
print("join      ", t(lambda: device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1)))
print("aggregate ", t(lambda: device.aggregate_hash(lineitem, Q1_GROUPBY, Q1_AGGREGATES, predicates=Q1_PREDICATES)))
r = device.table_scan(lineitem, SCAN_PREDICATE); print("scan.info ", t(lambda: r.info())); print("stats", t(lambda: device.last_stats()))
j = device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1); print("join.info ", t(lambda: j.info()))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    a = device.aggregate_hash(lineitem, Q1_GROUPBY, Q1_AGGREGATES, predicates=Q1_PREDICATES)
    s = device.table_scan(lineitem, SCAN_PREDICATE); s.free()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
