"""Compare what a rank receives through the NCCL exchange and through the fused P2P push (development aid, 2+ ranks)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from hyrise_b200 import capi, distributed as hd
from hyrise_b200.device import DeviceContext
from hyrise_b200.tpch import TpchTables, L_ORDERKEY, O_ORDERKEY
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); td = torch.device("cuda", lr)
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
tables = TpchTables(sf, first_order=rank * int(1_500_000 * sf))
device = DeviceContext(lr)
lineitem = device.upload(tables.lineitem); orders = device.upload(tables.orders)
lb = hd.chunk_bases(tables.lineitem.chunk_count, td)[rank]; ob = hd.chunk_bases(tables.orders.chunk_count, td)[rank]
sides = [hd.device_partition_side(device, orders, O_ORDERKEY, ob, world, td), hd.device_partition_side(device, lineitem, L_ORDERKEY, lb, world, td)]
(bk, br, bn), (pk, pr, pn) = hd.exchange_partitioned(sides)
torch.cuda.synchronize()
peers = hd.PeerExchange(device, td, capacity=2 * tables.lineitem.row_count + 65536)
bn2 = peers.push_side(orders, O_ORDERKEY, ob, 0); pn2 = peers.push_side(lineitem, L_ORDERKEY, lb, 2)
peers.barrier(); torch.cuda.synchronize()
bk2, br2 = peers.received(0, bn2); pk2, pr2 = peers.received(2, pn2)
for name, a, b, n, m in (("build keys", bk, bk2, bn, bn2), ("build rows", br, br2, bn, bn2), ("probe keys", pk, pk2, pn, pn2), ("probe rows", pr, pr2, pn, pn2)):
    a = a[:n].cpu().numpy(); b = b[:m].cpu().numpy()
    same = n == m and np.array_equal(a, b)
    print(f"rank {rank} {name}: counts {n} {m} equal {same}", flush=True)
    if not same and n == m:
        bad = np.nonzero(a != b)[0]
        print(f"   {len(bad)} mismatches, first at {bad[:8]}, last {bad[-3:]}; nccl {a[bad[:4]]} p2p {b[bad[:4]]}", flush=True)
        runs = np.split(bad, np.nonzero(np.diff(bad) != 1)[0] + 1)
        print(f"   {len(runs)} runs, lengths {[len(r) for r in runs[:12]]}, starts {[int(r[0]) for r in runs[:12]]}", flush=True)
peers.close(); dist.destroy_process_group()
