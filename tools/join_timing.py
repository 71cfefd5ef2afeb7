"""Wall / device / kernel time of repeated hyb_join_hash calls on the bench tables (development aid)."""
import sys, time
sys.path.insert(0, ".")
from bench import *  # noqa
from hyrise_b200.device import DeviceContext
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
tables = TpchTables(sf, seed=42)
device = DeviceContext(0)
lineitem = device.upload(tables.lineitem); orders = device.upload(tables.orders); device.synchronize()
for i in range(8):
    device.synchronize(); t0 = time.perf_counter()
    j = device.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, capi.JOIN_INNER, -1)
    t1 = time.perf_counter(); device.synchronize(); t2 = time.perf_counter()
    st = device.last_stats()
    print(f"call {1e3*(t1-t0):.2f} ms, +sync {1e3*(t2-t0):.2f} ms, device {st.device_ms:.2f}, kernels {st.dominant_kernel_ms:.2f}, pairs {j.info()[0]}", flush=True)
    j.free()
