"""Ad-hoc scan timing (development aid; bench.py is the contract)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from hyrise_b200 import capi
from hyrise_b200.device import DeviceContext, Predicate
from hyrise_b200.tpch import TpchTables, L_SHIPDATE

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t = time.time(); tp = TpchTables(sf, pinned=False); print("generate", time.time() - t, "s", tp.lineitem.row_count)
with DeviceContext(0) as dev:
    t = time.time(); li = dev.upload(tp.lineitem); dev.synchronize(); print("upload", time.time() - t, "s", tp.lineitem.host_bytes / 1e9, "GB")
    pred = Predicate(L_SHIPDATE, capi.PRED_LESS_THAN, "1995-01-01")
    times = []
    for i in range(iters):
        r = dev.table_scan(li, pred)
        st = dev.last_stats()
        times.append((st.dominant_kernel_ms, st.device_ms, st.algorithmic_bytes, st.output_rows))
        r.free()
    k = np.array([x[0] for x in times[3:]]); d = np.array([x[1] for x in times[3:]])
    b = times[-1][2]
    print(f"rows {tp.lineitem.row_count} matches {times[-1][3]} bytes {b/1e6:.1f} MB")
    print(f"kernel ms median {np.median(k):.4f} min {k.min():.4f}; op ms median {np.median(d):.4f}")
    print(f"achieved {b/np.median(k)/1e6:.1f} GB/s ({b/np.median(k)/1e6/6575.1:.3f} of measured 6575 GB/s)")
