"""Is pinned-host <-> device bandwidth sensitive to the NUMA node the pinned pages land on? (development aid)"""
import os, subprocess, sys, time
sys.path.insert(0, ".")
import torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
path = f"/sys/bus/pci/devices/{bus[4:].lower()}/numa_node"
node = open(path).read().strip() if os.path.exists(path) else "?"
print("gpu0 bus", bus, "numa node", node, "cpus now", len(os.sched_getaffinity(0)))
def measure(tag):
    host = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
    for direction in ("h2d", "d2h"):
        best = 0
        for _ in range(4):
            torch.cuda.synchronize(); t = time.perf_counter()
            (dev.copy_(host, non_blocking=True) if direction == "h2d" else host.copy_(dev, non_blocking=True)); torch.cuda.synchronize()
            best = max(best, (1 << 30) / (time.perf_counter() - t) / 1e9)
        print(f"{tag}: {direction} {best:.1f} GB/s", flush=True)
measure("default affinity")
if node not in ("?", "-1"):
    cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    cpus = set()
    for part in cpulist.split(","):
        lo, _, hi = part.partition("-"); cpus.update(range(int(lo), int(hi or lo) + 1))
    os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
    measure(f"bound to node {node} ({cpulist})")
    other = [n for n in os.listdir("/sys/devices/system/node") if n.startswith("node") and n != f"node{node}"]
    if other:
        cpulist = open(f"/sys/devices/system/node/{other[0]}/cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-"); cpus.update(range(int(lo), int(hi or lo) + 1))
        try:
            os.sched_setaffinity(0, cpus); measure(f"bound to {other[0]} ({cpulist})")
        except OSError as e:
            print("cannot bind to", other[0], e)
