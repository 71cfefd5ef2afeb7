"""Where does an end-to-end step (host blocks in, host results out) spend its wall time? (development aid)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from bench import *  # noqa
from hyrise_b200.device import DeviceContext, ROW_ID_DTYPE
tables = TpchTables(10.0, seed=42, pinned=True)
device = DeviceContext(0)
rows = tables.lineitem.row_count
scan_out = device.pinned_empty(rows, ROW_ID_DTYPE); jb = device.pinned_empty(rows, ROW_ID_DTYPE); jp = device.pinned_empty(rows, ROW_ID_DTYPE)
host_blocks = tables.host_blocks()
print("blocks", len(host_blocks), sum(b.bytes for b in host_blocks) / 1e6, "MB", flush=True)
def stamp(marks, name):
    device.synchronize(); marks.append((name, time.perf_counter()))
for step in range(5):
    marks = []; stamp(marks, "start")
    block_set = device.upload_blocks(host_blocks); stamp(marks, "upload_blocks")
    tl = device.upload_from_blocks(tables.lineitem, block_set); stamp(marks, "table lineitem")
    to = device.upload_from_blocks(tables.orders, block_set); stamp(marks, "table orders")
    scan = device.table_scan(tl, SCAN_PREDICATE); stamp(marks, "scan")
    matched = scan.to_host(scan_out); stamp(marks, "scan d2h")
    join = device.join_hash(to, O_ORDERKEY, tl, L_ORDERKEY, capi.JOIN_INNER, -1); stamp(marks, "join")
    pairs = join.to_host(jb, jp); stamp(marks, "join d2h")
    agg = device.aggregate_hash(tl, Q1_GROUPBY, Q1_AGGREGATES, predicates=Q1_PREDICATES); stamp(marks, "aggregate")
    scan.free(); join.free(); tl.drop(); to.drop(); device.free_blocks(block_set); stamp(marks, "free")
    print("step", step, " ".join(f"{name} {1e3 * (t - marks[i][1]):.1f}" for i, (name, t) in enumerate(marks[1:])),
          f"| total {1e3 * (marks[-1][1] - marks[0][1]):.1f} ms", flush=True)
