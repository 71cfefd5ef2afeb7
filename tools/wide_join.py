import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from hyrise_b200 import capi, distributed as hd
from hyrise_b200.device import DeviceContext
from hyrise_b200.tpch import TpchTables, L_ORDERKEY, O_ORDERKEY
tables = TpchTables(float(sys.argv[1]) if len(sys.argv) > 1 else 10.0)
device = DeviceContext(0); td = torch.device("cuda", 0)
lineitem = device.upload(tables.lineitem); orders = device.upload(tables.orders)
for i in range(8):
    t = time.time()
    pairs, offsets, b, p, result = hd.device_distributed_join(device, orders, O_ORDERKEY, lineitem, L_ORDERKEY, 8, 0, 0, td)
    st = device.last_stats(); device.synchronize()
    print("pairs", pairs, "kernel ms", st.dominant_kernel_ms, "op ms", st.device_ms, "wall", time.time() - t)
    result.free()
