import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from hyrise_b200 import capi, distributed as hd
from hyrise_b200.device import DeviceContext
from hyrise_b200.tpch import TpchTables, L_ORDERKEY, O_ORDERKEY
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); td = torch.device("cuda", lr)
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
tables = TpchTables(sf, first_order=rank * int(1_500_000 * sf))
device = DeviceContext(lr)
lineitem = device.upload(tables.lineitem); orders = device.upload(tables.orders)
lb = hd.chunk_bases(tables.lineitem.chunk_count, td)[rank]; ob = hd.chunk_bases(tables.orders.chunk_count, td)[rank]
def sync(): device.synchronize(); torch.cuda.synchronize()
peers = hd.PeerExchange(device, td, capacity=2 * tables.lineitem.row_count + 65536)
for i in range(6):
    sync(); t0 = time.time()
    bn = peers.push_side(orders, O_ORDERKEY, ob, 0); pn = peers.push_side(lineitem, L_ORDERKEY, lb, 2); sync(); t1 = time.time()
    peers.barrier(); sync(); t2 = time.time()
    bk, br = peers.received(0, bn); pk, pr = peers.received(2, pn)
    build = hd.DeviceTupleTable(device, bk, br, bn); probe = hd.DeviceTupleTable(device, pk, pr, pn); sync(); t3 = time.time()
    result = device.join_hash(build.table, 0, probe.table, 0, capi.JOIN_INNER, 8); st = device.last_stats(); sync(); t4 = time.time()
    if rank == 0:
        print(f"p2p: split+push {1e3*(t1-t0):.1f} barrier {1e3*(t2-t1):.1f} tables {1e3*(t3-t2):.1f} join wall {1e3*(t4-t3):.1f} "
              f"kernel {st.dominant_kernel_ms:.2f} pairs {result.info()[0]} build {build.count} probe {probe.count}", flush=True)
    result.free(); build.drop(); probe.drop()
for i in range(6):
    sync(); t0 = time.time()
    sides = [hd.device_partition_side(device, orders, O_ORDERKEY, ob, world, td),
             hd.device_partition_side(device, lineitem, L_ORDERKEY, lb, world, td)]; sync(); t1 = time.time()
    (bk, br, bn), (pk, pr, pn) = hd.exchange_partitioned(sides); sync(); t2 = time.time()
    build = hd.DeviceTupleTable(device, bk, br, bn); probe = hd.DeviceTupleTable(device, pk, pr, pn); sync(); t3 = time.time()
    result = device.join_hash(build.table, 0, probe.table, 0, capi.JOIN_INNER, 8); st = device.last_stats(); sync(); t4 = time.time()
    if rank == 0:
        print(f"partition {1e3*(t1-t0):.1f} exchange {1e3*(t2-t1):.1f} tables {1e3*(t3-t2):.1f} join wall {1e3*(t4-t3):.1f} "
              f"kernel {st.dominant_kernel_ms:.2f} op {st.device_ms:.2f} pairs {result.info()[0]} build {build.count} probe {probe.count}", flush=True)
    result.free(); build.drop(); probe.drop()
peers.close()
dist.destroy_process_group()
