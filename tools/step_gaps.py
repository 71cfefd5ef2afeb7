"""Wall time vs device time of the bench operators at a scale factor (development aid; GPU box).
    python tools/step_gaps.py [--sf 100]"""
import argparse
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import Q1_AGGREGATES, Q1_GROUPBY, Q1_PREDICATES, SCAN_PREDICATE  # noqa: E402
from hyrise_b200 import capi  # noqa: E402
from hyrise_b200.device import DeviceContext, build_scan_predicate  # noqa: E402
from hyrise_b200.tpch import L_ORDERKEY, O_ORDERKEY, TpchTables  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--sf", type=float, default=100.0)
    args = parser.parse_args()
    tables = TpchTables(args.sf, seed=42)
    device = DeviceContext(0)
    lineitem, orders = device.upload(tables.lineitem), device.upload(tables.orders)
    device.synchronize()

    def wall(name, run, repeat=7):
        walls, devices = [], []
        for _ in range(repeat):
            device.synchronize()
            begin = time.perf_counter()
            result = run()
            device.synchronize()
            walls.append((time.perf_counter() - begin) * 1e3)
            devices.append(device.last_stats().device_ms)
            if hasattr(result, "free"):
                result.free()
        print(f"{name:28s} wall {np.median(walls[1:]):8.3f} ms   device {np.median(devices[1:]):8.3f} ms   "
              f"gap {np.median(walls[1:]) - np.median(devices[1:]):7.3f} ms", flush=True)

    def host(name, run, repeat=7):
        samples = []
        for _ in range(repeat):
            begin = time.perf_counter()
            run()
            samples.append((time.perf_counter() - begin) * 1e3)
        print(f"{name:28s} host {np.median(samples[1:]):8.3f} ms", flush=True)

    wall("table_scan", lambda: device.table_scan(lineitem, SCAN_PREDICATE))
    wall("join_hash", lambda: device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1))
    wall("aggregate_hash", lambda: device.aggregate_hash(lineitem, Q1_GROUPBY, Q1_AGGREGATES, predicates=Q1_PREDICATES))
    host("build_scan_predicate (scan)", lambda: build_scan_predicate(tables.lineitem, SCAN_PREDICATE))
    host("build_scan_predicate (Q1)", lambda: build_scan_predicate(tables.lineitem, Q1_PREDICATES[0]))
    scan = device.table_scan(lineitem, SCAN_PREDICATE)
    join = device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1)
    host("scan.info", lambda: scan.info())
    host("join.info", lambda: join.info())
    host("last_stats", lambda: device.last_stats())
    begin = time.perf_counter()
    scan.free()
    join.free()
    print(f"free scan + join             host {(time.perf_counter() - begin) * 1e3:8.3f} ms")
    device.close()


if __name__ == "__main__":
    main()
