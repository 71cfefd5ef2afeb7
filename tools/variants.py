"""Kernel-variant experiments on the bench tables: env-selected variants, median kernel ms over repeated calls."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from bench import *  # noqa
from hyrise_b200.device import DeviceContext
tables = TpchTables(10.0, seed=42)
device = DeviceContext(0)
lineitem = device.upload(tables.lineitem); orders = device.upload(tables.orders); device.synchronize()
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
def run(f, n=12):
    xs = []
    for _ in range(n):
        flush.fill_(1); torch.cuda.synchronize()
        r = f(); device.synchronize(); xs.append(device.last_stats().dominant_kernel_ms)
        if hasattr(r, "free"): r.free()
    return float(np.median(xs[2:]))
for mb in ("2", "3", "4"):
    for pf in ("0",):
        os.environ["HYB_AGG_MIN_BLOCKS"] = mb; os.environ["HYB_AGG_PREFETCH"] = pf
        print(f"aggregate min_blocks {mb} prefetch {pf}: {run(lambda: device.aggregate_hash(lineitem, Q1_GROUPBY, Q1_AGGREGATES, predicates=Q1_PREDICATES)):.3f} ms", flush=True)
for pf in ("0", "1"):
    os.environ["HYB_SCAN_PREFETCH"] = pf
    print(f"scan prefetch {pf}: {run(lambda: device.table_scan(lineitem, SCAN_PREDICATE)):.4f} ms", flush=True)
print(f"join: {run(lambda: device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1)):.3f} ms", flush=True)
