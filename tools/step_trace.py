"""Host timeline of one bench step without extra synchronisation (development aid; GPU box).
    python tools/step_trace.py [--sf 25]"""
import argparse
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from bench import Q1_AGGREGATES, Q1_GROUPBY, Q1_PREDICATES, SCAN_PREDICATE  # noqa: E402
from hyrise_b200 import capi  # noqa: E402
from hyrise_b200.device import DeviceContext  # noqa: E402
from hyrise_b200.tpch import L_ORDERKEY, O_ORDERKEY, TpchTables  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--sf", type=float, default=25.0)
    args = parser.parse_args()
    tables = TpchTables(args.sf, seed=42)
    device = DeviceContext(0)
    lineitem, orders = device.upload(tables.lineitem), device.upload(tables.orders)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    stream_ptr = capi.C.c_void_p()
    capi.check(device.lib.hyb_context_stream(device.ptr, capi.C.byref(stream_ptr)))
    hyb_stream = torch.cuda.ExternalStream(stream_ptr.value, device="cuda:0")

    def flush_l2():
        with torch.cuda.stream(hyb_stream):
            flush.fill_(1)

    names = ["flush", "scan call", "flush", "join call", "flush", "aggregate call", "scan.info", "join.info", "frees"]
    samples = []
    for _ in range(12):
        device.synchronize()
        marks = [time.perf_counter()]
        flush_l2(); marks.append(time.perf_counter())
        scan = device.table_scan(lineitem, SCAN_PREDICATE); marks.append(time.perf_counter())
        flush_l2(); marks.append(time.perf_counter())
        join = device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1); marks.append(time.perf_counter())
        flush_l2(); marks.append(time.perf_counter())
        aggregate = device.aggregate_hash(lineitem, Q1_GROUPBY, Q1_AGGREGATES, predicates=Q1_PREDICATES); marks.append(time.perf_counter())
        scan.info(); marks.append(time.perf_counter())
        join.info(); marks.append(time.perf_counter())
        scan.free(); join.free(); marks.append(time.perf_counter())
        samples.append(np.diff(marks) * 1e3)
    median = np.median(np.array(samples[2:]), axis=0)
    for name, value in zip(names, median):
        print(f"{name:16s} {value:8.3f} ms")
    print(f"{'step':16s} {median.sum():8.3f} ms")
    # device times of the same operators
    for name, run in (("scan", lambda: device.table_scan(lineitem, SCAN_PREDICATE)),
                      ("join", lambda: device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1)),
                      ("aggregate", lambda: device.aggregate_hash(lineitem, Q1_GROUPBY, Q1_AGGREGATES, predicates=Q1_PREDICATES))):
        values = []
        for _ in range(5):
            result = run()
            stats = device.last_stats()
            values.append((stats.dominant_kernel_ms, stats.device_ms))
            if hasattr(result, "free"):
                result.free()
        print(f"{name:10s} kernel {np.median([v[0] for v in values]):7.3f} ms  device {np.median([v[1] for v in values]):7.3f} ms")
    device.close()


if __name__ == "__main__":
    main()
