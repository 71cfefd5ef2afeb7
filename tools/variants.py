"""Kernel time of the three bench operators under every option combination (development aid; run on the GPU box).
    python tools/variants.py [--sf 10]"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import Q1_AGGREGATES, Q1_GROUPBY, Q1_PREDICATES, SCAN_PREDICATE  # noqa: E402
from hyrise_b200 import capi  # noqa: E402
from hyrise_b200.device import DeviceContext  # noqa: E402
from hyrise_b200.tpch import L_ORDERKEY, O_ORDERKEY, TpchTables  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--sf", type=float, default=10.0)
    parser.add_argument("--repeat", type=int, default=5)
    parser.add_argument("--only", default="scan,join,aggregate")
    args = parser.parse_args()
    import torch
    tables = TpchTables(args.sf, seed=42)
    device = DeviceContext(0)
    lineitem, orders = device.upload(tables.lineitem), device.upload(tables.orders)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")

    def timed(name, options, run):
        for key, value in options.items():
            device.set_option(key, value)
        samples = []
        for _ in range(args.repeat + 1):
            flush.fill_(1)
            torch.cuda.synchronize()
            result = run()
            stats = device.last_stats()
            samples.append((stats.dominant_kernel_ms, stats.device_ms, stats.kernel_launches, stats.algorithmic_bytes))
            if hasattr(result, "free"):
                result.free()
        kernel = float(np.median([s[0] for s in samples[1:]]))
        print(f"{name:12s} {str(options):70s} kernel {kernel:.4f} ms  operator {np.median([s[1] for s in samples[1:]]):.4f} ms  "
              f"launches {samples[-1][2]}  {samples[-1][3] / kernel / 1e6:.0f} GB/s", flush=True)

    only = args.only.split(",")
    for options in ({"scan_two_pass": "1"}, {"scan_two_pass": "0", "scan_bulk": "1"}, {"scan_two_pass": "0", "scan_bulk": "0"}) \
            if "scan" in only else ():
        timed("scan", options, lambda: device.table_scan(lineitem, SCAN_PREDICATE))
    for options in ({"join_table": "auto", "join_span": "1", "join_rank": "ballot"},
                    {"join_table": "auto", "join_span": "1", "join_rank": "match"},
                    {"join_table": "direct", "join_span": "1", "join_rank": "ballot"},
                    {"join_table": "direct", "join_span": "0", "join_rank": "ballot"},
                    {"join_table": "hash", "join_span": "1", "join_rank": "ballot"}) if "join" in only else ():
        timed("join", options, lambda: device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1))
    device.set_option("join_table", "auto")
    for options in ({"aggregate_stream": "1", "aggregate_static_shapes": "1", "aggregate_stages": "4"},
                    {"aggregate_stream": "1", "aggregate_static_shapes": "1", "aggregate_stages": "5"},
                    {"aggregate_stream": "1", "aggregate_static_shapes": "1", "aggregate_stages": "6"},
                    {"aggregate_stream": "1", "aggregate_static_shapes": "1", "aggregate_stages": "7"},
                    {"aggregate_stream": "1", "aggregate_static_shapes": "1", "aggregate_stages": "8"},
                    {"aggregate_stream": "1", "aggregate_static_shapes": "0", "aggregate_stages": "6"},
                    {"aggregate_stream": "0"}) if "aggregate" in only else ():
        timed("aggregate", options, lambda: device.aggregate_hash(lineitem, Q1_GROUPBY, Q1_AGGREGATES, predicates=Q1_PREDICATES))
    device.close()


if __name__ == "__main__":
    main()
